// Developer microbenchmark: cycles of each softmax phase for ONE warp per SM sub-partition, no tensor-core traffic.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../video_styler_b200/csrc/ptx.cuh"
using namespace wvd;

template <int EMU_OF_4, int WARPS>
__global__ void phases(unsigned* out, float sl2, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = slot;
    const uint32_t s_tmem = tb + (warp / 4) * 128 + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t t_ld = 0, t_max = 0, t_exp = 0, t_st = 0;
    float m = 0.f, l = 0.f;
    for (int it = 0; it < iters; ++it) {
        uint32_t c0 = clock();
        uint32_t s[128];
        tmem_ld_32x32b_x32(s_tmem + 0, s + 0); tmem_ld_32x32b_x32(s_tmem + 32, s + 32);
        tmem_ld_32x32b_x32(s_tmem + 64, s + 64); tmem_ld_32x32b_x32(s_tmem + 96, s + 96);
        tc_wait_ld();
        uint32_t c1 = clock();
        float mx[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) mx[a] = fmaxf(__uint_as_float(s[2 * a]), __uint_as_float(s[2 * a + 1]));
#pragma unroll
        for (int c = 16; c < 128; c += 16)
#pragma unroll
            for (int a = 0; a < 8; ++a) mx[a] = fmax3(mx[a], __uint_as_float(s[c + 2 * a]), __uint_as_float(s[c + 2 * a + 1]));
        m = fmaxf(fmax3(fmax3(mx[0], mx[1], mx[2]), fmax3(mx[3], mx[4], mx[5]), fmaxf(mx[6], mx[7])), m);
        uint32_t c2 = clock();
        const float neg_m = -m * sl2;
        const uint64_t sl2_2 = f2_pack(sl2, sl2), negm_2 = f2_pack(neg_m, neg_m);
        uint64_t la = f2_pack(0.f, 0.f), lb = f2_pack(0.f, 0.f);
        uint32_t pk[64];
#pragma unroll
        for (int c = 0; c < 128; c += 2) {
            const uint64_t x2 = f2_fma(f2_pack(__uint_as_float(s[c]), __uint_as_float(s[c + 1])), sl2_2, negm_2);
            float p0, p1;
            if (((c >> 1) & 3) < EMU_OF_4) exp2_poly2(x2, p0, p1);
            else { float x0, x1; f2_unpack(x2, x0, x1); p0 = fast_exp2(x0); p1 = fast_exp2(x1); }
            if (c & 2) lb = f2_add(lb, f2_pack(p0, p1)); else la = f2_add(la, f2_pack(p0, p1));
            pk[c >> 1] = pack_bf16x2(p0, p1);
        }
        float a0, a1; f2_unpack(f2_add(la, lb), a0, a1); l += a0 + a1;
        uint32_t c3 = clock();
        tmem_st_32x32b_x32(s_tmem + 0, pk + 0); tmem_st_32x32b_x32(s_tmem + 32, pk + 32);
        tc_wait_st();
        uint32_t c4 = clock();
        t_ld += c1 - c0; t_max += c2 - c1; t_exp += c3 - c2; t_st += c4 - c3;
    }
    if (lane == 0 && warp == 0) { out[0] = t_ld / iters; out[1] = t_max / iters; out[2] = t_exp / iters; out[3] = t_st / iters; out[4] = __float_as_uint(l + m); }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

template <int E, int W> void run() {
    unsigned* d; cudaMalloc(&d, 64);
    phases<E, W><<<1, W * 32>>>(d, 0.1275f, 200);
    cudaDeviceSynchronize();
    unsigned h[8]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("emu %d/4, %d warps (%d per SMSP): ld %u  max %u  exp %u  st %u  total %u cycles (%s)\n", E, W, W / 4, h[0], h[1], h[2], h[3],
           h[0] + h[1] + h[2] + h[3], cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}
int main() { run<0, 4>(); run<1, 4>(); run<2, 4>(); run<3, 4>(); run<0, 8>(); run<1, 8>(); run<2, 8>(); return 0; }
