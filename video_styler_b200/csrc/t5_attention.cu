// K11: umT5 encoder self-attention (head_dim 64) with additive relative-position bias and key mask.
//
//   out[i, h] = softmax_j( q_h[i].k_h[j] * scale + bias[h][j - i] , masked keys -> finfo.min ) v_h
//   (replaces T5Attention.forward, diffsynth/models/wan_video_text_encoder.py:55-90: "T5 does not use scaling", fp32
//    softmax, bias = T5RelativeEmbedding(lq, lk) :141-153 with 32 bidirectional buckets)
//
// Not a tensor-core kernel on purpose: one prompt is 512 tokens x 64 heads x 64 dims = 4.3 GFLOP of attention per layer
// (0.1 TFLOP per prompt, twice per video, against 175 PFLOP of DiT work per video); it is latency-bound on any
// implementation, the projections around it (98 % of the encoder's FLOPs) run on the tcgen05 GEMM.  What matters here is
// exactness: the reference materialises bf16 scores, adds a bf16 bias, soft-maxes in fp32, rounds the probabilities to
// bf16 and contracts with V -- the kernel keeps every one of those rounding points (two passes over the keys: row
// maximum / sum first, then normalised probabilities), so the bf16 result matches the reference to accumulation order.
//
// One CTA = one head x 32 query rows, 4 warps x 8 queries; keys in chunks of 64 through shared memory (K as fp32 rows
// padded to 68 words so that 128-bit loads are conflict-free, Q rows read as 128-bit broadcasts, V in the I/O dtype).  Pass A: lane l scores keys l and l + 32 of the chunk for the warp's 8
// queries; running max / sum per query.  Pass B: the same scores again, p = bf16(exp(s - m) / sum) to shared memory, lane
// l accumulates output dims 2l, 2l + 1.
#include <float.h>
#include <math.h>

#include "host_utils.h"
#include "ptx.cuh"

namespace wvd {
namespace t5 {

constexpr int HD = 64, QB = 32, KB = 64, WARPS = 4, QPW = QB / WARPS;

template <typename T> struct Num;
template <> struct Num<__nv_bfloat16> {
    static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
    static __device__ __forceinline__ void ld2(const __nv_bfloat16* p, float& a, float& b) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
        a = f.x; b = f.y;
    }
    static __device__ __forceinline__ float rnd(float v) { return bf16_round(v); }
    static __device__ __forceinline__ float lowest() { return -3.3895313892515355e38f; }     // torch.finfo(bfloat16).min
};
template <> struct Num<float> {
    static __device__ __forceinline__ float ld(const float* p) { return *p; }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
    static __device__ __forceinline__ void ld2(const float* p, float& a, float& b) {
        const float2 f = *reinterpret_cast<const float2*>(p);
        a = f.x; b = f.y;
    }
    static __device__ __forceinline__ float rnd(float v) { return v; }
    static __device__ __forceinline__ float lowest() { return -FLT_MAX; }                       // torch.finfo(float32).min
};

template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
t5_attention_kernel(const T* __restrict__ q, long long ldq, const T* __restrict__ k, long long ldk, const T* __restrict__ v,
                    long long ldv, const T* __restrict__ bias, long long ld_bias, const int* __restrict__ key_mask,
                    T* __restrict__ out, long long ldo, int lq, int lk, float scale) {
    __shared__ __align__(16) float q_s[QB][HD];
    __shared__ __align__(16) float k_s[KB][HD + 4];      // 68-word rows: 128-bit loads of 8 consecutive lanes hit 32 distinct banks
    __shared__ __align__(16) T v_s[KB][HD];
    float* p_s = &k_s[0][0];      // pass B: the probabilities reuse the K chunk once every warp has its scores (8 KB of 16.6)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.y;
    const int q0 = blockIdx.x * QB;
    const T* qh = q + head * HD;
    const T* kh = k + head * HD;
    const T* vh = v + head * HD;
    // bias row of this head, indexed by (j - i) + (lq - 1)
    const T* bh = bias != nullptr ? bias + static_cast<long long>(head) * ld_bias + (lq - 1) : nullptr;

    for (int e = threadIdx.x; e < QB * HD; e += WARPS * 32) {
        const int r = e / HD, c = e % HD;
        q_s[r][c] = (q0 + r < lq) ? Num<T>::ld(qh + static_cast<long long>(q0 + r) * ldq + c) : 0.f;
    }
    const int qw = warp * QPW;                 // first query of this warp within the CTA tile
    float m[QPW], l[QPW];
#pragma unroll
    for (int i = 0; i < QPW; ++i) { m[i] = -INFINITY; l[i] = 0.f; }
    float acc[QPW][2];
#pragma unroll
    for (int i = 0; i < QPW; ++i) acc[i][0] = acc[i][1] = 0.f;

    // scores of the warp's 8 queries against keys (lane, lane + 32) of the chunk in shared memory
    auto scores = [&](int j0, float (&s)[QPW][2]) {
#pragma unroll
        for (int i = 0; i < QPW; ++i) s[i][0] = s[i][1] = 0.f;
#pragma unroll 4
        for (int d = 0; d < HD; d += 4) {
            const float4 k0 = *reinterpret_cast<const float4*>(&k_s[lane][d]);
            const float4 k1 = *reinterpret_cast<const float4*>(&k_s[lane + 32][d]);
#pragma unroll
            for (int i = 0; i < QPW; ++i) {
                const float4 qv = *reinterpret_cast<const float4*>(&q_s[qw + i][d]);      // broadcast
                s[i][0] = fmaf(qv.w, k0.w, fmaf(qv.z, k0.z, fmaf(qv.y, k0.y, fmaf(qv.x, k0.x, s[i][0]))));
                s[i][1] = fmaf(qv.w, k1.w, fmaf(qv.z, k1.z, fmaf(qv.y, k1.y, fmaf(qv.x, k1.x, s[i][1]))));
            }
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int j = j0 + lane + 32 * t;
            const bool in_range = j < lk;
            const bool masked = in_range && key_mask != nullptr && key_mask[j] == 0;
#pragma unroll
            for (int i = 0; i < QPW; ++i) {
                const int qi = q0 + qw + i;
                float x = Num<T>::rnd(s[i][t] * scale);                       // the einsum's output dtype
                float b = 0.f;
                if (bh != nullptr && in_range && qi < lq) b = Num<T>::ld(bh + (j - qi));
                if (masked) b = Num<T>::lowest();                             // masked_fill_ on the bias tensor
                x = Num<T>::rnd(x + b);
                s[i][t] = in_range ? x : -INFINITY;                           // keys past the end do not exist
            }
        }
    };
    auto load_chunk = [&](int j0, bool with_v) {
        __syncthreads();
        for (int e = threadIdx.x; e < KB * HD; e += WARPS * 32) {
            const int r = e / HD, c = e % HD;
            const bool ok = j0 + r < lk;
            k_s[r][c] = ok ? Num<T>::ld(kh + static_cast<long long>(j0 + r) * ldk + c) : 0.f;
            if (with_v) v_s[r][c] = ok ? vh[static_cast<long long>(j0 + r) * ldv + c] : T(0.f);
        }
        __syncthreads();
    };

    // ---- pass A: row maximum and sum of exponentials ----
    for (int j0 = 0; j0 < lk; j0 += KB) {
        load_chunk(j0, false);
        float s[QPW][2];
        scores(j0, s);
#pragma unroll
        for (int i = 0; i < QPW; ++i) {
            float mx = fmaxf(s[i][0], s[i][1]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float m_new = fmaxf(m[i], mx);
            float e = expf(s[i][0] - m_new) + expf(s[i][1] - m_new);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
            l[i] = l[i] * expf(m[i] - m_new) + e;
            m[i] = m_new;
        }
    }

    // ---- pass B: normalised probabilities (rounded like the reference's .type_as) times V ----
    for (int j0 = 0; j0 < lk; j0 += KB) {
        load_chunk(j0, true);
        float s[QPW][2];
        scores(j0, s);
        __syncthreads();                       // every warp is done reading K: its space now holds P
        float* pw = p_s + warp * QPW * KB;
#pragma unroll
        for (int i = 0; i < QPW; ++i) {
            pw[i * KB + lane] = Num<T>::rnd(__fdiv_rn(expf(s[i][0] - m[i]), l[i]));       // exp / sum like torch.softmax
            pw[i * KB + lane + 32] = Num<T>::rnd(__fdiv_rn(expf(s[i][1] - m[i]), l[i]));
        }
        __syncwarp();
#pragma unroll 2
        for (int j = 0; j < KB; j += 4) {
            float v0[4], v1[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) Num<T>::ld2(&v_s[j + t][2 * lane], v0[t], v1[t]);
#pragma unroll
            for (int i = 0; i < QPW; ++i) {
                const float4 pv = *reinterpret_cast<const float4*>(&pw[i * KB + j]);       // broadcast
                acc[i][0] = fmaf(pv.w, v0[3], fmaf(pv.z, v0[2], fmaf(pv.y, v0[1], fmaf(pv.x, v0[0], acc[i][0]))));
                acc[i][1] = fmaf(pv.w, v1[3], fmaf(pv.z, v1[2], fmaf(pv.y, v1[1], fmaf(pv.x, v1[0], acc[i][1]))));
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < QPW; ++i) {
        const int qi = q0 + qw + i;
        if (qi < lq) {
            T* o = out + static_cast<long long>(qi) * ldo + head * HD + 2 * lane;
            Num<T>::st(o, acc[i][0]);
            Num<T>::st(o + 1, acc[i][1]);
        }
    }
}

}  // namespace t5
}  // namespace wvd

extern "C" __attribute__((visibility("default"))) int wvd_attention_bias_fwd(
    const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* bias, int64_t ld_bias,
    const int* key_mask, void* out, int64_t ldo, int num_heads, int64_t lq, int64_t lk, int head_dim, float scale, int dtype,
    wvd_stream_t stream) {
    using namespace wvd;
    WVD_REQUIRE(dtype == WVD_BF16 || dtype == WVD_F32, "wvd_attention_bias_fwd: bad dtype %d", dtype);
    WVD_REQUIRE(head_dim == t5::HD, "wvd_attention_bias_fwd: head_dim must be 64 (got %d)", head_dim);
    WVD_REQUIRE(q && k && v && out, "wvd_attention_bias_fwd: null pointer");
    WVD_REQUIRE(num_heads > 0 && num_heads <= 65535 && lq > 0 && lk > 0 && lq < (1ll << 24) && lk < (1ll << 24),
                "wvd_attention_bias_fwd: bad sizes heads=%d lq=%lld lk=%lld", num_heads, (long long)lq, (long long)lk);
    const int64_t width = (int64_t)num_heads * head_dim;
    WVD_REQUIRE(ldq >= width && ldk >= width && ldv >= width && ldo >= width, "wvd_attention_bias_fwd: leading dims must cover heads*64");
    WVD_REQUIRE(bias == nullptr || ld_bias >= lq + lk - 1, "wvd_attention_bias_fwd: the bias table needs lq + lk - 1 entries per head");
    dim3 grid((unsigned)((lq + t5::QB - 1) / t5::QB), (unsigned)num_heads);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == WVD_BF16)
        t5::t5_attention_kernel<__nv_bfloat16><<<grid, t5::WARPS * 32, 0, st>>>(
            (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv, (const __nv_bfloat16*)bias,
            ld_bias, key_mask, (__nv_bfloat16*)out, ldo, (int)lq, (int)lk, scale);
    else
        t5::t5_attention_kernel<float><<<grid, t5::WARPS * 32, 0, st>>>((const float*)q, ldq, (const float*)k, ldk, (const float*)v,
                                                                         ldv, (const float*)bias, ld_bias, key_mask, (float*)out,
                                                                         ldo, (int)lq, (int)lk, scale);
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}
