// fp32 mode of the parity contract (BASELINE.json: "fp32 mode must be within 1e-4"): plain CUDA-core kernels with
// fp32 FMA accumulation (no TF32, no tensor cores).  They exist so that the whole operator surface can be checked
// against the fp32 CPU oracle at config c1; they are not the tuned path and are never selected for bf16 tensors.
#include <math.h>

#include "host_utils.h"
#include "ptx.cuh"

namespace wvd {
namespace f32 {

constexpr int TM = 64, TN = 64, TK = 16;

__device__ __forceinline__ float gelu_tanh_f32(float x) {
    const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
    return 0.5f * x * (1.0f + tanhf(u));
}

// C[M,N] = epi(A[M,K] . W[N,K]^T + bias); 64x64 tile, 256 threads, 4x4 micro-tile per thread
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ W, long long ldw,
                const float* __restrict__ bias, float* __restrict__ C, long long ldc, int M, int N, int K, int epi,
                const float* __restrict__ gate, const float* __restrict__ res, long long ldr) {
    __shared__ float As[TK][TM + 1];
    __shared__ float Ws[TK][TN + 1];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += TK) {
        for (int i = threadIdx.x; i < TM * TK; i += 256) {
            const int r = i / TK, c = i % TK;
            const int gm = m0 + r, gn = n0 + r, gk = k0 + c;
            As[c][r] = (gm < M && gk < K) ? A[static_cast<long long>(gm) * lda + gk] : 0.f;
            Ws[c][r] = (gn < N && gk < K) ? W[static_cast<long long>(gn) * ldw + gk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Ws[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float y = acc[i][j] + (bias ? bias[gn] : 0.f);
            if (epi == WVD_EPI_BIAS_GELU) y = gelu_tanh_f32(y);
            if (epi == WVD_EPI_BIAS_GELU_T5) y = 0.5f * y * (1.0f + tanhf(0.7978845608028654f * (y + 0.044715f * (y * y * y))));
            if (epi == WVD_EPI_BIAS_GATE_RES) y = gate[gn] * y;
            if (epi == WVD_EPI_BIAS_RES || epi == WVD_EPI_BIAS_GATE_RES) y = res[static_cast<long long>(gm) * ldr + gn] + y;
            if (epi == WVD_EPI_BIAS_MUL) y = y * res[static_cast<long long>(gm) * ldr + gn];
            C[static_cast<long long>(gm) * ldc + gn] = y;
        }
    }
}

// one warp per (query row, head); each lane owns 4 of the 128 channels; online softmax over the keys
__global__ void __launch_bounds__(256)
attention_f32_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ k, long long ldk,
                     const float* __restrict__ v, long long ldv, float* __restrict__ out, long long ldo, int heads,
                     long long sq, long long sk, float scale) {
    const int lane = threadIdx.x & 31;
    const long long w = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (w >= sq * heads) return;
    const long long row = w / heads;
    const int h = static_cast<int>(w % heads);
    const float4 qv = *reinterpret_cast<const float4*>(q + row * ldq + h * 128 + lane * 4);
    float m = -INFINITY, l = 0.f;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long j = 0; j < sk; ++j) {
        const float4 kv = *reinterpret_cast<const float4*>(k + j * ldk + h * 128 + lane * 4);
        float d = qv.x * kv.x + qv.y * kv.y + qv.z * kv.z + qv.w * kv.w;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) d += __shfl_xor_sync(0xffffffffu, d, s);
        d *= scale;
        const float m_new = fmaxf(m, d);
        const float alpha = __expf(m - m_new);
        const float pj = __expf(d - m_new);
        const float4 vv = *reinterpret_cast<const float4*>(v + j * ldv + h * 128 + lane * 4);
        o.x = o.x * alpha + pj * vv.x;
        o.y = o.y * alpha + pj * vv.y;
        o.z = o.z * alpha + pj * vv.z;
        o.w = o.w * alpha + pj * vv.w;
        l = l * alpha + pj;
        m = m_new;
    }
    const float inv = 1.f / l;
    *reinterpret_cast<float4*>(out + row * ldo + h * 128 + lane * 4) = make_float4(o.x * inv, o.y * inv, o.z * inv, o.w * inv);
}

}  // namespace f32
}  // namespace wvd

using namespace wvd;

extern "C" __attribute__((visibility("default"))) int wvd_gemm_f32(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* C,
                            int64_t ldc, int64_t M, int64_t N, int64_t K, int epilogue, const void* gate,
                            const void* residual, int64_t ldr, wvd_stream_t stream) {
    WVD_REQUIRE(A && W && C && M > 0 && N > 0 && K > 0, "wvd_gemm_f32: bad arguments");
    WVD_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "wvd_gemm_f32: dimension too large");
    WVD_REQUIRE(epilogue >= WVD_EPI_BIAS && epilogue <= WVD_EPI_BIAS_GELU_T5, "wvd_gemm_f32: bad epilogue %d", epilogue);
    if (epilogue == WVD_EPI_BIAS_RES || epilogue == WVD_EPI_BIAS_GATE_RES || epilogue == WVD_EPI_BIAS_MUL) WVD_REQUIRE(residual, "wvd_gemm_f32: residual missing");
    if (epilogue == WVD_EPI_BIAS_GATE_RES) WVD_REQUIRE(gate, "wvd_gemm_f32: gate missing");
    dim3 grid((unsigned)((N + f32::TN - 1) / f32::TN), (unsigned)((M + f32::TM - 1) / f32::TM));
    f32::gemm_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)A, lda, (const float*)W, ldw, (const float*)bias,
                                                                  (float*)C, ldc, (int)M, (int)N, (int)K, epilogue,
                                                                  (const float*)gate, (const float*)residual, ldr);
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}

extern "C" __attribute__((visibility("default"))) int wvd_attention_fwd_f32(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                     void* out, int64_t ldo, int num_heads, int64_t sq, int64_t sk, int head_dim,
                                     float scale, wvd_stream_t stream) {
    WVD_REQUIRE(q && k && v && out && sq > 0 && sk > 0 && num_heads > 0, "wvd_attention_fwd_f32: bad arguments");
    WVD_REQUIRE(head_dim == 128, "wvd_attention_fwd_f32: head_dim must be 128 (got %d)", head_dim);
    WVD_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0, "wvd_attention_fwd_f32: ld must be multiples of 4");
    const long long warps = (long long)sq * num_heads;
    f32::attention_f32_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        (const float*)q, ldq, (const float*)k, ldk, (const float*)v, ldv, (float*)out, ldo, num_heads, sq, sk, scale);
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}
