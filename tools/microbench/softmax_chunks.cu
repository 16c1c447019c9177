// Developer microbenchmark: the softmax step of the attention kernel for ONE warp per SM sub-partition, no
// tensor-core traffic.  Variants isolate the cost of each piece of the in-kernel step:
//   VAR 0  plain 128-column exp2 loop, one TMEM store at the end
//   VAR 1  kernel structure: chunk 0 with row max, async store, chunk 1 head, handoff 0, chunk 1 tail, handoff 1
//   VAR 2  VAR 1 without the row max
//   VAR 3  VAR 1 without the mid handoff (store only)
//   VAR 4  VAR 1 without vote/branch (no redo loop)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../video_styler_b200/csrc/softmax_math.cuh"
using namespace wvd;
using namespace wvd::attn;

template <int VAR>
__global__ void __launch_bounds__(384, 1) phases(unsigned* out, float sl2, int iters) {
    __shared__ uint32_t slot;
    __shared__ uint64_t bar[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar[0]), 4); mbar_init(smem_u32(&bar[1]), 4); fence_barrier_init(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = slot;
    if (warp >= 4) {
        setmaxnreg_dec<72>();
    } else {
        setmaxnreg_inc<216>();
        const uint32_t s_tmem = tb + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        {   // fill S with scores in [-40, 0] / sl2
            uint32_t v[32];
            for (int c = 0; c < 4; ++c) {
                for (int e = 0; e < 32; ++e) {
                    uint32_t h = (threadIdx.x * 131 + c * 32 + e) * 2654435761u;
                    v[e] = __float_as_uint(-(float)(h >> 8) * (40.0f / 16777216.0f) / sl2);
                }
                tmem_st_32x32b_x32(s_tmem + 256 + c * 32, v);
            }
            tc_wait_st();
        }
        uint32_t t_ld = 0, t_a = 0, t_b = 0;
        float m = 0.f, l = 0.f;
        const uint64_t sl2_2 = f2_pack(sl2, sl2);
        for (int it = 0; it < iters; ++it) {
            uint32_t c0 = clock();
            uint32_t s[128];
            tmem_ld_32x32b_x32(s_tmem + 256, s + 0); tmem_ld_32x32b_x32(s_tmem + 288, s + 32);
            tmem_ld_32x32b_x32(s_tmem + 320, s + 64); tmem_ld_32x32b_x32(s_tmem + 352, s + 96);
            tc_wait_ld();
            uint32_t c1 = clock(), c2, c3;
            if (VAR == 0) {
                uint32_t pk[64];
                const float neg_m = -m * sl2;
                const float la = exp_chunk<0, 128, 0>(s, pk, sl2_2, f2_pack(neg_m, neg_m));
                l += la;
                c2 = clock();
                store_p<64>(s_tmem, pk);
                tc_wait_st();
                c3 = clock();
            } else {
                float la, lb;
                uint32_t pka[32], pkb[32];
                {
                    bool redo = false;
                    float mx = 0.f;
#pragma unroll 1
                    for (;;) {
                        if (redo) m = mx;
                        const float neg_m = -m * sl2;
                        float mxa[4] = {m, m, m, m};
                        if (VAR == 2) la = exp_chunk<0, 64, 0, false>(s, pka, sl2_2, f2_pack(neg_m, neg_m));
                        else la = exp_chunk<0, 64, 0, true>(s, pka, sl2_2, f2_pack(neg_m, neg_m), mxa);
                        mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
                        if (VAR == 4) break;
                        if (redo || !__any_sync(0xffffffffu, (mx - m) * sl2 > 8.0f)) break;
                        redo = true;
                    }
                    store_p<32>(s_tmem, pka);
                }
                c2 = clock();
                {
                    const float neg_m = -m * sl2;
                    const uint64_t negm_2 = f2_pack(neg_m, neg_m);
                    lb = exp_chunk<64, 96, 0>(s, pkb, sl2_2, negm_2);
                    if (VAR != 3) {
                        tc_wait_st(); tc_fence_before(); __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&bar[0]));
                    }
                    lb += exp_chunk<96, 128, 0>(s, pkb + 16, sl2_2, negm_2);
                    store_p<32>(s_tmem + 32, pkb);
                    tc_wait_st(); tc_fence_before(); __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bar[1]));
                }
                l += la + lb;
                c3 = clock();
            }
            t_ld += c1 - c0; t_a += c2 - c1; t_b += c3 - c2;
        }
        if (lane == 0 && warp == 0) { out[0] = t_ld / iters; out[1] = t_a / iters; out[2] = t_b / iters; out[3] = __float_as_uint(l + m); }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

template <int VAR> void run() {
    unsigned* d; cudaMalloc(&d, 64);
    phases<VAR><<<1, 384>>>(d, 0.1275f, 200);
    cudaDeviceSynchronize();
    unsigned h[8]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("var %d: ld %u  phaseA %u  phaseB %u  total %u cycles (%s)\n", VAR, h[0], h[1], h[2],
           h[0] + h[1] + h[2], cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}
int main() {
    run<0>(); run<1>(); run<2>(); run<3>(); run<4>();
    return 0;
}
