// Developer microbenchmark: cycles per warp-instruction of the ops used in the attention softmax (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define UNROLL 16

__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }

template <int OP>
__global__ void bench(float* out, unsigned* cyc, float a, float b) {
    float x[UNROLL];
    uint64_t y[UNROLL];
    for (int i = 0; i < UNROLL; ++i) { x[i] = a + i + threadIdx.x; y[i] = f2_pack(a + i, b + i); }
    const uint64_t a2 = f2_pack(a, a), b2 = f2_pack(b, b);
    __syncthreads();
    unsigned t0 = clock();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
            if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[i]) : "l"(a2), "l"(b2));
            if (OP == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(y[i]) : "l"(b2));
            if (OP == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
            if (OP == 5) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i]), "f"(a)); x[i] = __uint_as_float(r); }
            if (OP == 6) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
            if (OP == 7) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(x[i]) : "f"(a));
            if (OP == 8) { int r; asm volatile("mad.lo.s32 %0, %1, 8388608, %2;" : "=r"(r) : "r"(__float_as_int(x[i])), "r"(__float_as_int(a))); x[i] = __int_as_float(r); }
            if (OP == 9) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a));
        }
    }
    unsigned t1 = clock();
    float s = 0;
    for (int i = 0; i < UNROLL; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(y[i])); s += x[i] + lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name) {
    float* out; unsigned* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4);
    for (int warps : {4, 8, 16}) {
        bench<OP><<<1, warps * 32>>>(out, cyc, 1.0001f, 0.5f);
        cudaDeviceSynchronize();
        unsigned c; cudaMemcpy(&c, cyc, 4, cudaMemcpyDeviceToHost);
        double per = (double)c / (ITERS * UNROLL);
        printf("%-22s warps/SM %2d (per SMSP %d): %.2f cyc per warp-instr per warp -> %.2f cyc/instr per SMSP\n", name, warps, warps / 4, per, per / (warps / 4));
    }
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("FFMA (3 reg)");
    run<7>("FFMA (imm c)");
    run<6>("FADD");
    run<1>("FFMA2 (f32x2)");
    run<2>("FADD2 (f32x2)");
    run<3>("MUFU.EX2");
    run<4>("FMNMX3");
    run<9>("FMNMX");
    run<5>("F2FP bf16x2");
    run<8>("IMAD (shift-add)");
    return 0;
}
