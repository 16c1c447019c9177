// Developer microbenchmark: does instruction class B slow down a concurrent MUFU.EX2 stream on the same scheduler?
// 8 warps per SM (2 per SMSP): warps 0-3 run op A, warps 4-7 run op B, both for the same number of instructions.
// Prints cycles per instruction of each group; alone-runs give the baseline.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
#define UNROLL 16
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }

template <int OP>
__device__ __forceinline__ void body(float (&x)[UNROLL], uint64_t (&y)[UNROLL], float a, float b, uint64_t a2, uint64_t b2) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
        if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        if (OP == 1) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
        if (OP == 2) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i]), "f"(a)); x[i] = __uint_as_float(r); }
        if (OP == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[i]) : "l"(a2), "l"(b2));
        if (OP == 4) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(y[i]) : "l"(b2));
        if (OP == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a));
        if (OP == 6) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
        if (OP == 7) { int r; asm volatile("mad.lo.s32 %0, %1, 8388608, %2;" : "=r"(r) : "r"(__float_as_int(x[i])), "r"(__float_as_int(a))); x[i] = __int_as_float(r); }
    }
}

template <int OPA, int OPB>
__global__ void bench(float* out, unsigned* cyc, float a, float b, int nb) {
    float x[UNROLL]; uint64_t y[UNROLL];
    for (int i = 0; i < UNROLL; ++i) { x[i] = a + i + threadIdx.x; y[i] = f2_pack(a + i, b + i); }
    const uint64_t a2 = f2_pack(a, a), b2 = f2_pack(b, b);
    const int grp = threadIdx.x >> 7;
    __syncthreads();
    unsigned t0 = clock();
    if (grp == 0) { for (int it = 0; it < ITERS; ++it) body<OPA>(x, y, a, b, a2, b2); }
    else { for (int it = 0; it < nb; ++it) body<OPB>(x, y, a, b, a2, b2); }
    unsigned t1 = clock();
    float s = 0;
    for (int i = 0; i < UNROLL; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(y[i])); s += x[i] + lo + hi; }
    out[threadIdx.x] = s;
    if ((threadIdx.x & 127) == 0) cyc[grp] = t1 - t0;
}
template <int OPA, int OPB> void run(const char* na, const char* nb_, int nb) {
    float* out; unsigned* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    bench<OPA, OPB><<<1, 256>>>(out, cyc, 1.0001f, 0.5f, nb);
    cudaDeviceSynchronize();
    unsigned c[2]; cudaMemcpy(c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("A=%-10s alongside B=%-10s x%-5d: A %.2f cyc/instr   B %.2f cyc/instr\n", na, nb_, nb, (double)c[0] / (ITERS * UNROLL), nb ? (double)c[1] / ((double)nb * UNROLL) : 0.0);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0, 0>("MUFU", "(none)", 0);
    run<0, 1>("MUFU", "FMNMX3", 4 * ITERS);
    run<0, 5>("MUFU", "FMNMX", 8 * ITERS);
    run<0, 2>("MUFU", "F2FP", 4 * ITERS);
    run<0, 3>("MUFU", "FFMA2", 4 * ITERS);
    run<0, 4>("MUFU", "FADD2", 4 * ITERS);
    run<0, 6>("MUFU", "FFMA", 8 * ITERS);
    run<0, 7>("MUFU", "IMAD", 8 * ITERS);
    run<1, 0>("FMNMX3", "MUFU", ITERS / 4);
    run<2, 0>("F2FP", "MUFU", ITERS / 4);
    run<1, 2>("FMNMX3", "F2FP", ITERS);
    run<3, 4>("FFMA2", "FADD2", ITERS);
    run<3, 1>("FFMA2", "FMNMX3", ITERS);
    return 0;
}
