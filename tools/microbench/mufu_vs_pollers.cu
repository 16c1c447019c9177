// Developer microbenchmark: does a warp spinning on mbarrier.try_wait (what every waiting role of a warp-specialised
// kernel does) slow down a MUFU-bound softmax stream on the same scheduler?
//   4 worker warps (one per SM sub-partition) run the 128-column exp2 loop of the attention softmax;
//   P poller warps per sub-partition spin on an mbarrier that never completes (MODE 1: plain try_wait loop,
//   MODE 2: try_wait with a 2 us suspend hint = NANOSLEEP form, MODE 3: blocked on a named barrier = no issue at all).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../video_styler_b200/csrc/softmax_math.cuh"
using namespace wvd;
using namespace wvd::attn;

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(unsigned* out, float sl2, int iters, int pollers_per_smsp) {
    __shared__ uint64_t bar;
    __shared__ volatile int done;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); done = 0; }
    __syncthreads();
    if (warp < 4) {
        uint32_t s[128];
        for (int e = 0; e < 128; ++e) s[e] = __float_as_uint(-(float)((threadIdx.x * 131 + e) % 97) * 0.25f);
        float l = 0.f;
        const uint64_t sl2_2 = f2_pack(sl2, sl2);
        uint32_t t0 = clock();
        for (int it = 0; it < iters; ++it) {
            uint32_t pk[64];
            const float neg_m = -l * 1e-9f;
            l += exp_chunk<128, 0, 128, 0>(s, pk, sl2_2, f2_pack(neg_m, neg_m));
            for (int e = 0; e < 64; e += 16) s[e] ^= pk[e] & 1u;       // keep pk alive
        }
        uint32_t t1 = clock();
        if (lane == 0 && warp == 0) { out[0] = (t1 - t0) / iters; out[1] = __float_as_uint(l); }
        __syncwarp();
        if (threadIdx.x == 0) done = 1;
        if (MODE == 3) { asm volatile("bar.arrive 1, %0;" ::"r"(32 + 32 * 4 * pollers_per_smsp)); }
    } else if (warp < 4 + 4 * pollers_per_smsp) {
        if (MODE == 1) { while (!done) { mbar_try_wait(smem_u32(&bar), 0); } }
        if (MODE == 2) { while (!done) { mbar_try_wait_hint(smem_u32(&bar), 0, 2000); } }
        if (MODE == 3) { asm volatile("bar.sync 1, %0;" ::"r"(32 + 32 * 4 * pollers_per_smsp)); }
    }
}

template <int MODE> void run(int pollers) {
    unsigned* d; cudaMalloc(&d, 64);
    k<MODE><<<1, 512>>>(d, 0.1275f, 200, pollers);
    cudaDeviceSynchronize();
    unsigned h[2]; cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("mode %d (%s), %d poller warp(s) per sub-partition: %u cycles per 128-column exp2 row (%s)\n", MODE,
           MODE == 1 ? "try_wait spin" : MODE == 2 ? "try_wait + suspend hint" : "named barrier", pollers, h[0],
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}
int main() {
    run<3>(0); run<3>(1); run<3>(3);
    run<1>(1); run<1>(2); run<1>(3);
    run<2>(1); run<2>(3);
    return 0;
}
