// K1-K4 and the unfused residual forms: HBM-bound, vectorised (16-byte) kernels, one warp per token row.
// A row (<= 6144 bf16 channels) is read once into registers, reduced with warp shuffles (no block barrier),
// and written once: algorithmic traffic = 1 read + 1 write per element.
//
//   wvd_ln_modulate      LayerNorm (+ AdaLN modulate or affine)         wan_video_dit.py:64-65,206-208,225-228,262-269
//   wvd_qk_rmsnorm_rope  full-width RMSNorm(q,k) * weight (+ 3-D RoPE)  wan_video_dit.py:92-111,141-145,177-178
//   wvd_scale_add        x + y*scale (VACE hint injection)             wan_video_new.py:1445-1450
//   wvd_gate_residual    x + gate*y                                    wan_video_dit.py:189-194
//   wvd_ulysses_*        all-to-all send/receive layouts               distributed/xdit_context_parallel.py:110-131
//
// In bf16 mode every intermediate rounding of the reference's eager bf16 expressions is reproduced
// (see the comments at each site), so the kernels are drop-in at the bit level up to fp32 reduction order.
#include "host_utils.h"
#include "ptx.cuh"

namespace wvd {
namespace ew {

template <typename T> struct VecIO;

template <> struct VecIO<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* f) {
        const uint4 u = *reinterpret_cast<const uint4*>(p);
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
    }
    static __device__ __forceinline__ uint4 load_raw(const __nv_bfloat16* p) {
        uint4 u;   // streaming read: the row is consumed once, keep it out of L1
        asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
        return u;
    }
    static __device__ __forceinline__ void unpack(const uint4& u, float* f) {
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* f) {
        uint4 u;
        u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
        u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
        *reinterpret_cast<uint4*>(p) = u;
    }
    static __device__ __forceinline__ float rnd(float x) { return bf16_round(x); }
    static __device__ __forceinline__ float ld1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

template <> struct VecIO<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void load(const float* p, float* f) {
        const float4 u = *reinterpret_cast<const float4*>(p);
        f[0] = u.x; f[1] = u.y; f[2] = u.z; f[3] = u.w;
    }
    static __device__ __forceinline__ uint4 load_raw(const float* p) {
        uint4 u;
        asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
        return u;
    }
    static __device__ __forceinline__ void unpack(const uint4& u, float* f) {
        f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
    }
    static __device__ __forceinline__ void store(float* p, const float* f) {
        *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    }
    static __device__ __forceinline__ float rnd(float x) { return x; }
    static __device__ __forceinline__ float ld1(const float* p) { return *p; }
    static __device__ __forceinline__ void st1(float* p, float v) { *p = v; }
};

// Opaque to the optimiser: stops it from keeping the unpacked fp32 copy of a packed row alive across passes
// (which would quadruple the live registers and defeat the occupancy the packed layout buys).
template <int MAXV>
__device__ __forceinline__ void keep_packed(uint4 (&raw)[MAXV]) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        asm volatile("" : "+r"(raw[i].x), "+r"(raw[i].y), "+r"(raw[i].z), "+r"(raw[i].w));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int WARPS_PER_BLOCK = 8;

// ------------------------------------------------------------------------------------------------
// LayerNorm + modulate / affine
// ------------------------------------------------------------------------------------------------
// The row stays PACKED in registers (16-byte vectors, MAXV per lane) and is unpacked on the fly in each of the three
// passes (sum, squared deviations, output): 4x fewer live registers than an fp32 copy, so 2 blocks of 8 warps fit per
// SM and every lane keeps MAXV independent 16-byte loads in flight.
template <typename T, int MAXV, bool MODULATE, bool AFFINE>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, (MAXV <= 20 ? 2 : 1))
ln_modulate_kernel(const T* __restrict__ x, long long ldx, const T* __restrict__ shift, const T* __restrict__ scale,
                   const T* __restrict__ weight, const T* __restrict__ bias, T* __restrict__ out, long long ldo,
                   long long n_tokens, int dim, float eps) {
    using IO = VecIO<T>;
    constexpr int VE = IO::N;
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (row >= n_tokens) return;
    const int nvec = dim / VE;
    const T* xr = x + row * ldx;
    uint4 raw[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nvec) raw[i] = IO::load_raw(xr + vi * VE);
    }
    // Two statistics passes over the packed registers (mean, then squared deviations about the mean): the same
    // cancellation-free fp32 arithmetic as F.layer_norm; unpacking twice is cheaper than keeping an fp32 copy live.
    float s1 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        if (lane + 32 * i < nvec) {
            float f[VE];
            IO::unpack(raw[i], f);
#pragma unroll
            for (int e = 0; e < VE; ++e) s1 += f[e];
        }
    }
    const float mean = warp_sum(s1) / dim;
    keep_packed(raw);
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        if (lane + 32 * i < nvec) {
            float f[VE];
            IO::unpack(raw[i], f);
#pragma unroll
            for (int e = 0; e < VE; ++e) { const float d = f[e] - mean; s2 = fmaf(d, d, s2); }
        }
    }
    const float rstd = rsqrtf(warp_sum(s2) / dim + eps);
    keep_packed(raw);
    T* orow = out + row * ldo;
    if constexpr (MODULATE && !AFFINE && sizeof(T) == 2) {
        // bf16 fast path of modulate(LN(x), shift, scale): the reference evaluates x * (1 + scale) + shift in bf16, one
        // rounding per operation (wan_video_dit.py:64-65) -- exactly what the packed bf16x2 instructions do (HADD2 /
        // HMUL2 round to nearest even once), so the whole tail runs on 2 elements per instruction instead of
        // emulating each rounding in fp32.  LN output: (x - mean) * rstd in fp32, rounded to bf16 by the pack.
        const __nv_bfloat162 one2 = __floats2bfloat162_rn(1.0f, 1.0f);
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int vi = lane + 32 * i;
            if (vi < nvec) {
                float f[VE];
                IO::unpack(raw[i], f);
                const uint4 sc4 = __ldg(reinterpret_cast<const uint4*>(scale + vi * VE));
                const uint4 sh4 = __ldg(reinterpret_cast<const uint4*>(shift + vi * VE));
                const uint32_t scw[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, shw[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
                uint32_t ow[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const __nv_bfloat162 ln = __floats2bfloat162_rn((f[2 * q] - mean) * rstd, (f[2 * q + 1] - mean) * rstd);
                    const __nv_bfloat162 sc2 = *reinterpret_cast<const __nv_bfloat162*>(&scw[q]);
                    const __nv_bfloat162 sh2 = *reinterpret_cast<const __nv_bfloat162*>(&shw[q]);
                    const __nv_bfloat162 r = __hadd2_rn(__hmul2_rn(ln, __hadd2_rn(one2, sc2)), sh2);   // _rn: never contracted into an FMA
                    ow[q] = *reinterpret_cast<const uint32_t*>(&r);
                }
                *reinterpret_cast<uint4*>(orow + vi * VE) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nvec) {
            float o[VE], f[VE];
            IO::unpack(raw[i], f);
            if (AFFINE) {
                // F.layer_norm(x, weight, bias): ((x - mean) * rstd) * w + b in fp32, one rounding
                float w[VE], b[VE];
                IO::load(weight + vi * VE, w);
                IO::load(bias + vi * VE, b);
#pragma unroll
                for (int e = 0; e < VE; ++e) o[e] = (f[e] - mean) * rstd * w[e] + b[e];
            } else {
#pragma unroll
                for (int e = 0; e < VE; ++e) o[e] = IO::rnd((f[e] - mean) * rstd);       // LN output -> dtype
            }
            if (MODULATE) {
                // modulate(): x * (1 + scale) + shift evaluated in the tensor dtype (wan_video_dit.py:64-65)
                float sc[VE], sh[VE];
                IO::load(scale + vi * VE, sc);
                IO::load(shift + vi * VE, sh);
#pragma unroll
                for (int e = 0; e < VE; ++e) {
                    if (AFFINE) o[e] = IO::rnd(o[e]);
                    const float one_plus = IO::rnd(1.0f + sc[e]);
                    o[e] = IO::rnd(o[e] * one_plus) + sh[e];
                }
            }
            IO::store(orow + vi * VE, o);
        }
    }
}

template <typename T, int MAXV>
int launch_ln(const void* x, int64_t ldx, const void* shift, const void* scale, const void* weight, const void* bias,
              void* out, int64_t ldo, int64_t n, int dim, float eps, cudaStream_t s) {
    const unsigned grid = static_cast<unsigned>((n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    const dim3 block(WARPS_PER_BLOCK * 32);
    const bool mod = shift != nullptr, aff = weight != nullptr;
#define WVD_LN_LAUNCH(M, A)                                                                                      \
    ln_modulate_kernel<T, MAXV, M, A><<<grid, block, 0, s>>>((const T*)x, ldx, (const T*)shift, (const T*)scale, \
                                                             (const T*)weight, (const T*)bias, (T*)out, ldo, n, dim, eps)
    if (mod && aff) WVD_LN_LAUNCH(true, true);
    else if (mod) WVD_LN_LAUNCH(true, false);
    else if (aff) WVD_LN_LAUNCH(false, true);
    else WVD_LN_LAUNCH(false, false);
#undef WVD_LN_LAUNCH
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}

// ------------------------------------------------------------------------------------------------
// RMSNorm(q), RMSNorm(k) over the full hidden dim, * weight, then 3-D RoPE on adjacent channel pairs
// ------------------------------------------------------------------------------------------------
// `dst(vi)` = where vector vi (channels [vi*VE, vi*VE + VE)) of the finished row goes: the row itself, or -- Ulysses,
// fused exchange -- the receive buffer of the rank that owns the vector's head (a peer pointer: the store travels NVLink)
template <typename T, int MAXV, bool ROPE, typename Dst>
__device__ __forceinline__ void rms_rope_row(const T* __restrict__ xr, const T* __restrict__ w, Dst dst,
                                             int dim, float eps, int lane, const float2* cs /*[VE/2]*/) {
    using IO = VecIO<T>;
    constexpr int VE = IO::N;
    const int nvec = dim / VE;
    uint4 raw[MAXV];                     // packed row (see ln_modulate_kernel)
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nvec) raw[i] = IO::load_raw(xr + vi * VE);
    }
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        if (lane + 32 * i < nvec) {
            float f[VE];
            IO::unpack(raw[i], f);
#pragma unroll
            for (int e = 0; e < VE; ++e) sq += f[e] * f[e];
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / dim + eps);
    keep_packed(raw);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nvec) {
            float o[VE], f[VE];
            IO::unpack(raw[i], f);
            if constexpr (sizeof(T) == 2) {
                // bf16: norm(x.float()).to(dtype) * weight (wan_video_dit.py:109-111) = one pack (rounds the norm) and
                // one packed bf16x2 multiply (rounds the product) per two elements -- the reference's rounding points
                const uint4 w4 = __ldg(reinterpret_cast<const uint4*>(w + vi * VE));
                const uint32_t wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const __nv_bfloat162 n2 = __floats2bfloat162_rn(f[2 * q] * rstd, f[2 * q + 1] * rstd);
                    const __nv_bfloat162 m2 = __hmul2_rn(n2, *reinterpret_cast<const __nv_bfloat162*>(&wv[q]));
                    const float2 mf = __bfloat1622float2(m2);
                    o[2 * q] = mf.x;
                    o[2 * q + 1] = mf.y;
                }
            } else {
                float ww[VE];
                IO::load(w + vi * VE, ww);
#pragma unroll
                for (int e = 0; e < VE; ++e)   // norm(x.float()).to(dtype) * weight  (wan_video_dit.py:109-111)
                    o[e] = IO::rnd(IO::rnd(f[e] * rstd) * ww[e]);
            }
            if (ROPE) {
#pragma unroll
                for (int pr = 0; pr < VE / 2; ++pr) {   // (re, im) * (cos + i sin)  (wan_video_dit.py:92-97)
                    const float re = o[2 * pr], im = o[2 * pr + 1];
                    o[2 * pr] = re * cs[pr].x - im * cs[pr].y;
                    o[2 * pr + 1] = re * cs[pr].y + im * cs[pr].x;
                }
            }
            IO::store(dst(vi), o);
        }
    }
}

struct PeerPtrs { uint4* p[WVD_MAX_PEERS]; };

// blockIdx.y selects the tensor (0 = q, 1 = k): q and k rows are independent, one warp each.
template <typename T, int MAXV, bool ROPE>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, (MAXV <= 20 ? 2 : 1))
qk_rmsnorm_rope_kernel(const T* __restrict__ q, long long ldq, const T* __restrict__ k, long long ldk,
                       const T* __restrict__ wq, const T* __restrict__ wk, T* __restrict__ qo, long long ldqo,
                       T* __restrict__ ko, long long ldko, long long n_tokens, int dim, float eps,
                       const float2* __restrict__ rope_cs, const int* __restrict__ frame_ids, int gf, int gh, int gw,
                       long long token_offset) {
    constexpr int VE = VecIO<T>::N;
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (row >= n_tokens) return;
    float2 cs[VE / 2];
    if (ROPE) {
        // this lane always owns the same pairs of every head: channel (lane*VE) % 128 (+ 0..VE-1)
        const int pair0 = ((lane * VE) % 128) / 2;
        if (gf == 0) {
            // per-token mode: rope_cs is (n_tokens, 64) (cos, sin) -- the reference's `freqs` tensor (N, 1, 64) complex
            // as DiTBlock.forward receives it (wan_video_dit.py:214-230), cast to fp32 pairs by the binding
#pragma unroll
            for (int pr = 0; pr < VE / 2; ++pr) cs[pr] = __ldg(rope_cs + row * 64 + pair0 + pr);
        } else {
            const long long n = token_offset + row;
            const int pw = static_cast<int>(n % gw);
            const int ph = static_cast<int>((n / gw) % gh);
            // rows past the grid are the zero padding of the last Ulysses shard (wan_video_new.py:1414-1416): they are
            // never attended nor returned, any valid table entry will do
            int pf = min(static_cast<int>(n / (static_cast<long long>(gw) * gh)), gf - 1);
            if (frame_ids != nullptr) pf = frame_ids[pf];
#pragma unroll
            for (int pr = 0; pr < VE / 2; ++pr) {
                const int j = pair0 + pr;
                int axis, jj, pos;
                if (j < 22) { axis = 0; jj = j; pos = pf; }
                else if (j < 43) { axis = 1; jj = j - 22; pos = ph; }
                else { axis = 2; jj = j - 43; pos = pw; }
                cs[pr] = __ldg(rope_cs + (static_cast<long long>(axis) * 1024 + pos) * 32 + jj);
            }
        }
    }
    if (blockIdx.y == 0) {
        T* orow = qo + row * ldqo;
        rms_rope_row<T, MAXV, ROPE>(q + row * ldq, wq, [orow](int vi) { return orow + vi * VE; }, dim, eps, lane, cs);
    } else {
        T* orow = ko + row * ldko;
        rms_rope_row<T, MAXV, ROPE>(k + row * ldk, wk, [orow](int vi) { return orow + vi * VE; }, dim, eps, lane, cs);
    }
}

// The same kernel with the Ulysses all-to-all of q and k FUSED into its stores (bf16, RoPE on): the finished vector of
// head h goes straight into the receive buffer of rank h / heads_local, at
//   recv[dest][(rank * n_local + row)][which][h % heads_local][channel % 128]      (which: 0 = q, 1 = k; v is slot 2)
// -- the layout the attention kernel reads in place (see ulysses_scatter_kernel).  No separate pack / scatter pass for q, k.
template <int MAXV>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, (MAXV <= 20 ? 2 : 1))
qk_rmsnorm_rope_scatter_kernel(const __nv_bfloat16* __restrict__ q, long long ldq, const __nv_bfloat16* __restrict__ k,
                               long long ldk, const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wk,
                               PeerPtrs recv, long long n_local, int dim, float eps, const float2* __restrict__ rope_cs,
                               const int* __restrict__ frame_ids, int gf, int gh, int gw, long long token_offset,
                               int heads_local, int rank) {
    using T = __nv_bfloat16;
    constexpr int VE = 8;
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (row >= n_local) return;
    float2 cs[VE / 2];
    {
        const int pair0 = ((lane * VE) % 128) / 2;
        const long long n = token_offset + row;
        const int pw = static_cast<int>(n % gw);
        const int ph = static_cast<int>((n / gw) % gh);
        int pf = min(static_cast<int>(n / (static_cast<long long>(gw) * gh)), gf - 1);      // pad rows of the last shard
        if (frame_ids != nullptr) pf = frame_ids[pf];
#pragma unroll
        for (int pr = 0; pr < VE / 2; ++pr) {
            const int j = pair0 + pr;
            int axis, jj, pos;
            if (j < 22) { axis = 0; jj = j; pos = pf; }
            else if (j < 43) { axis = 1; jj = j - 22; pos = ph; }
            else { axis = 2; jj = j - 43; pos = pw; }
            cs[pr] = __ldg(rope_cs + (static_cast<long long>(axis) * 1024 + pos) * 32 + jj);
        }
    }
    const int which = blockIdx.y;
    const long long per_row = 3ll * heads_local * 16;                   // 16-byte vectors per receive-buffer row
    const long long row_base = (static_cast<long long>(rank) * n_local + row) * per_row + static_cast<long long>(which) * heads_local * 16;
    auto dst = [&](int vi) {
        const int head = vi >> 4;                                       // 16 vectors of 8 channels per 128-channel head
        const int dest = head / heads_local;
        return reinterpret_cast<T*>(recv.p[dest] + row_base + (head - dest * heads_local) * 16 + (vi & 15));
    };
    if (which == 0) rms_rope_row<T, MAXV, true>(q + row * ldq, wq, dst, dim, eps, lane, cs);
    else rms_rope_row<T, MAXV, true>(k + row * ldk, wk, dst, dim, eps, lane, cs);
}

template <typename T, int MAXV>
int launch_rms(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* wq, const void* wk, void* qo,
               int64_t ldqo, void* ko, int64_t ldko, int64_t n, int dim, float eps, const void* rope_cs,
               const int32_t* frame_ids, int gf, int gh, int gw, int64_t token_offset, cudaStream_t s) {
    const dim3 grid(static_cast<unsigned>((n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK), k != nullptr ? 2 : 1);
    const dim3 block(WARPS_PER_BLOCK * 32);
    if (rope_cs != nullptr)
        qk_rmsnorm_rope_kernel<T, MAXV, true><<<grid, block, 0, s>>>(
            (const T*)q, ldq, (const T*)k, ldk, (const T*)wq, (const T*)wk, (T*)qo, ldqo, (T*)ko, ldko, n, dim, eps,
            (const float2*)rope_cs, frame_ids, gf, gh, gw, token_offset);
    else
        qk_rmsnorm_rope_kernel<T, MAXV, false><<<grid, block, 0, s>>>(
            (const T*)q, ldq, (const T*)k, ldk, (const T*)wq, (const T*)wk, (T*)qo, ldqo, (T*)ko, ldko, n, dim, eps,
            nullptr, nullptr, 1, 1, 1, 0);
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}

// ------------------------------------------------------------------------------------------------
// residual forms
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
scale_add_kernel(const T* __restrict__ x, const T* __restrict__ y, float scale, T* __restrict__ out, long long nvec) {
    using IO = VecIO<T>;
    constexpr int VE = IO::N;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float a[VE], b[VE];
        IO::load(x + i * VE, a);
        IO::load(y + i * VE, b);
#pragma unroll
        for (int e = 0; e < VE; ++e) a[e] = __fadd_rn(a[e], IO::rnd(__fmul_rn(b[e], scale)));   // x + hint * vace_scale (two roundings, no FMA)
        IO::store(out + i * VE, a);
    }
}

// CFG combine + Euler step of the denoising loop in one pass (wan_video_new.py:535,540; flow_match.py:72-82), with the
// rounding points of the reference's tensor-dtype expressions:
//   noise = v_nega + cfg * (v_posi - v_nega)        (each operation rounded to the tensor dtype; skipped if v_nega == NULL)
//   out   = x + noise * dsigma                      (product rounded, then the sum)
template <typename T>
__global__ void __launch_bounds__(256)
cfg_euler_kernel(const T* __restrict__ x, const T* __restrict__ vp, const T* __restrict__ vn, float cfg, float dsigma,
                 T* __restrict__ out, long long nvec) {
    using IO = VecIO<T>;
    constexpr int VE = IO::N;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float a[VE], p[VE], n[VE];
        IO::load(x + i * VE, a);
        IO::load(vp + i * VE, p);
        if (vn != nullptr) {
            IO::load(vn + i * VE, n);
#pragma unroll
            for (int e = 0; e < VE; ++e) {
                const float d = IO::rnd(__fsub_rn(p[e], n[e]));
                p[e] = IO::rnd(__fadd_rn(n[e], IO::rnd(__fmul_rn(cfg, d))));
            }
        }
#pragma unroll
        for (int e = 0; e < VE; ++e) a[e] = __fadd_rn(a[e], IO::rnd(__fmul_rn(p[e], dsigma)));
        IO::store(out + i * VE, a);
    }
}

// ------------------------------------------------------------------------------------------------
// Keyframe editor step (diffsynth/pipelines/wan_video_editor.py:107-165, 362-390): CFG combine of the joint
// (main | edited keyframes) velocity, velocity-field correction at the keyframe positions, Euler update of BOTH latent
// sets -- one pass, with the reference's per-operation rounding in the tensor dtype:
//   v      = v_nega + cfg * (v_posi - v_nega)                               (:368, three roundings; skipped without v_nega)
//   z_diff = z_main[key_k] - z_edit[k] ; v_diff = v_main[key_k] - v_edit[k]
//   r_k    = z_diff - v_diff * dt ; corr = alpha * r_k                      (:143-150)
//   v_main[key_k] += corr ; v_edit[k] -= beta * corr  (beta > 0 only)       (:153-160)
//   euler != 0: out = z + v * dsigma (flow_match.py:72-82) ; euler == 0: out = the corrected velocity
// Velocities come as (BC, frames, HW) slabs with their own BC strides, so the concatenated (B, C, T+K, H, W) model
// output is read in place (main part at offset 0, keyframe part at offset T*HW) and so are separate tensors.
// frame_to_key[t] = k if main frame t is keyframe k, else -1; key_idx[k] = its main frame.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
editor_step_kernel(const T* __restrict__ z_main, const T* __restrict__ z_edit, const T* __restrict__ vp_main,
                   const T* __restrict__ vp_edit, const T* __restrict__ vn_main, const T* __restrict__ vn_edit,
                   long long v_main_bc_stride, long long v_edit_bc_stride, const int* __restrict__ frame_to_key,
                   const int* __restrict__ key_idx, int bc, int t_frames, int k_frames, long long hw, float cfg, float dt,
                   float alpha, float beta, float dsigma, int euler, T* __restrict__ out_main, T* __restrict__ out_edit) {
    using IO = VecIO<T>;
    constexpr int VE = IO::N;
    const long long hwv = hw / VE;
    const long long per_bc = static_cast<long long>(t_frames + k_frames) * hwv;
    const long long total = per_bc * bc;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / per_bc);
        const long long rem = i - b * per_bc;
        const int f = static_cast<int>(rem / hwv);
        const long long x = (rem - f * hwv) * VE;
        const bool is_main = f < t_frames;
        const int k = is_main ? frame_to_key[f] : f - t_frames;          // keyframe slot, -1: an ordinary main frame
        const int t = is_main ? f : key_idx[k];                          // main frame
        const long long zm_off = (static_cast<long long>(b) * t_frames + t) * hw + x;
        const long long vm_off = b * v_main_bc_stride + static_cast<long long>(t) * hw + x;
        float vm[VE], ve[VE], zm[VE], ze[VE], n[VE];
        auto velocity = [&](const T* vp, const T* vn, long long off, float* v) {
            IO::load(vp + off, v);
            if (vn != nullptr) {
                IO::load(vn + off, n);
#pragma unroll
                for (int e = 0; e < VE; ++e) {
                    const float d = IO::rnd(__fsub_rn(v[e], n[e]));
                    v[e] = IO::rnd(__fadd_rn(n[e], IO::rnd(__fmul_rn(cfg, d))));
                }
            }
        };
        velocity(vp_main, vn_main, vm_off, vm);
        IO::load(z_main + zm_off, zm);
        if (k >= 0) {
            const long long ze_off = (static_cast<long long>(b) * k_frames + k) * hw + x;
            velocity(vp_edit, vn_edit, b * v_edit_bc_stride + static_cast<long long>(k) * hw + x, ve);
            IO::load(z_edit + ze_off, ze);
#pragma unroll
            for (int e = 0; e < VE; ++e) {
                const float zd = IO::rnd(__fsub_rn(zm[e], ze[e]));
                const float vd = IO::rnd(__fsub_rn(vm[e], ve[e]));
                const float rk = IO::rnd(__fsub_rn(zd, IO::rnd(__fmul_rn(vd, dt))));
                const float corr = IO::rnd(__fmul_rn(alpha, rk));
                vm[e] = IO::rnd(__fadd_rn(vm[e], corr));
                if (beta > 0.f) ve[e] = IO::rnd(__fsub_rn(ve[e], IO::rnd(__fmul_rn(beta, corr))));
            }
            if (!is_main) {
                if (euler) {
#pragma unroll
                    for (int e = 0; e < VE; ++e) ve[e] = __fadd_rn(ze[e], IO::rnd(__fmul_rn(ve[e], dsigma)));
                }
                IO::store(out_edit + ze_off, ve);
                continue;
            }
        }
        if (euler) {
#pragma unroll
            for (int e = 0; e < VE; ++e) vm[e] = __fadd_rn(zm[e], IO::rnd(__fmul_rn(vm[e], dsigma)));
        }
        IO::store(out_main + zm_off, vm);
    }
}

// ------------------------------------------------------------------------------------------------
// VAE tile blending (WanVideoVAE.tiled_decode / tiled_encode, diffsynth/models/wan_video_vae.py:1081-1204):
//   values[:, :, :, y0:y0+th, x0:x0+tw] += tile * mask ; weight[..., same window] += mask ; finally values / weight (clamped
//   for decode).  mask[y][x] = min(ramp_h(y), ramp_w(x)), ramp = 1 inside, (i + 1) / border over the first `border`
//   pixels of a side that is not a volume boundary (the right / bottom ramp wins where the two overlap, like the
//   reference's two slice assignments), evaluated in fp32 and rounded to the tensor dtype; products and sums round per
//   operation like the reference's tensor ops.  The reference keeps `values` / `weight` on the CPU and ships every tile
//   over PCIe; here both stay in HBM.  The weight is the same for every channel and frame: one (H, W) plane.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tile_ramp(int i, int n, bool lo_bound, bool hi_bound, int border) {
    const int r = n - 1 - i;
    if (!hi_bound && r < border) return __fdiv_rn(static_cast<float>(r + 1), static_cast<float>(border));
    if (!lo_bound && i < border) return __fdiv_rn(static_cast<float>(i + 1), static_cast<float>(border));
    return 1.0f;
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
tile_blend_kernel(T* __restrict__ values, T* __restrict__ weight, const T* __restrict__ tile, int planes, int H, int W, int th,
                  int tw, int y0, int x0, int bounds, int border_h, int border_w) {
    using IO = VecIO<T>;
    // 32-bit index arithmetic (the launcher guarantees planes * th * tw < 2^31): 64-bit div / mod per element cost more
    // than the memory traffic of this kernel
    const unsigned twv = tw / V;
    const unsigned total = static_cast<unsigned>(planes) * th * twv;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned rest = i / twv;
        const int xv = static_cast<int>(i - rest * twv);
        const int pl = static_cast<int>(rest / th);               // channel * frames + frame
        const int y = static_cast<int>(rest - pl * th);
        const int x = xv * V;
        const float mh = tile_ramp(y, th, bounds & 1, bounds & 2, border_h);
        const long long src = (static_cast<long long>(pl) * th + y) * tw + x;
        const long long dst = (static_cast<long long>(pl) * H + y0 + y) * W + x0 + x;
        const long long wdst = static_cast<long long>(y0 + y) * W + x0 + x;
        float t[V], a[V], w[V], m[V];
#pragma unroll
        for (int e = 0; e < V; ++e) m[e] = IO::rnd(fminf(mh, tile_ramp(x + e, tw, bounds & 4, bounds & 8, border_w)));
        if (V == 1) {
            t[0] = IO::ld1(tile + src); a[0] = IO::ld1(values + dst);
            IO::st1(values + dst, __fadd_rn(a[0], IO::rnd(__fmul_rn(t[0], m[0]))));
            if (pl == 0) IO::st1(weight + wdst, __fadd_rn(IO::ld1(weight + wdst), m[0]));
        } else {
            IO::load(tile + src, t);
            IO::load(values + dst, a);
#pragma unroll
            for (int e = 0; e < V; ++e) a[e] = __fadd_rn(a[e], IO::rnd(__fmul_rn(t[e], m[e])));
            IO::store(values + dst, a);
            if (pl == 0) {
                IO::load(weight + wdst, w);
#pragma unroll
                for (int e = 0; e < V; ++e) w[e] = __fadd_rn(w[e], m[e]);
                IO::store(weight + wdst, w);
            }
        }
    }
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
tile_finalize_kernel(T* __restrict__ values, const T* __restrict__ weight, int planes, long long hw, int clamp, float lo, float hi) {
    using IO = VecIO<T>;
    const long long hwv = hw / V;
    const long long total = planes * hwv;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long p = (i % hwv) * V;
        float a[V], w[V];
        if (V == 1) { a[0] = IO::ld1(values + i); w[0] = IO::ld1(weight + p); }
        else { IO::load(values + i * V, a); IO::load(weight + p, w); }
#pragma unroll
        for (int e = 0; e < V; ++e) {
            a[e] = IO::rnd(__fdiv_rn(a[e], w[e]));
            if (clamp) a[e] = fminf(fmaxf(a[e], lo), hi);
        }
        if (V == 1) IO::st1(values + i, a[0]);
        else IO::store(values + i * V, a);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
gate_residual_kernel(const T* __restrict__ x, const T* __restrict__ gate, const T* __restrict__ y, T* __restrict__ out,
                     long long n_tokens, int dim) {
    using IO = VecIO<T>;
    constexpr int VE = IO::N;
    const int vpr = dim / VE;
    const long long nvec = n_tokens * vpr;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % vpr);
        float a[VE], b[VE], g[VE];
        IO::load(x + i * VE, a);
        IO::load(y + i * VE, b);
        IO::load(gate + c * VE, g);
#pragma unroll
        for (int e = 0; e < VE; ++e) a[e] = __fadd_rn(a[e], IO::rnd(__fmul_rn(g[e], b[e])));
        IO::store(out + i * VE, a);
    }
}

// ------------------------------------------------------------------------------------------------
// Ulysses layouts (bf16, 16-byte vectors)
// ------------------------------------------------------------------------------------------------
// send[dest][row][which][hl][c] = qkv[row][which*H*hd + (dest*Hl + hl)*hd + c]
__global__ void __launch_bounds__(256)
ulysses_pack_kernel(const uint4* __restrict__ qkv, long long ld_vec, uint4* __restrict__ send, long long n_local,
                    int heads, int hd_vec, int world) {
    const int hl_n = heads / world;
    const long long per_row = 3ll * hl_n * hd_vec;                  // vectors per (dest,row)
    const long long total = static_cast<long long>(world) * n_local * per_row;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % hd_vec);
        long long t = i / hd_vec;
        const int hl = static_cast<int>(t % hl_n); t /= hl_n;
        const int which = static_cast<int>(t % 3); t /= 3;
        const long long row = t % n_local;
        const int dest = static_cast<int>(t / n_local);
        const long long src = row * ld_vec + (static_cast<long long>(which) * heads + dest * hl_n + hl) * hd_vec + c;
        send[i] = qkv[src];
    }
}

// Fused pack + all-to-all: the same gather as ulysses_pack_kernel, but each destination's chunk is stored straight into
// THAT rank's receive buffer over NVLink (peer pointers from a symmetric-memory rendezvous), in the layout the attention
// kernel reads: recv_dest[(rank*n_local + row)][which][hl][c].  1,280-byte contiguous runs per (dest,row,which) at P = 8.
// which0: first of the three tensors to move (0 = q, k and v; 2 = v only, when q and k were scattered by the RoPE kernel)
__global__ void __launch_bounds__(256)
ulysses_scatter_kernel(const uint4* __restrict__ qkv, long long ld_vec, PeerPtrs recv, long long n_local, int heads,
                       int hd_vec, int world, int rank, int which0) {
    const int hl_n = heads / world;
    const int nw = 3 - which0;
    const long long per_row = 3ll * hl_n * hd_vec;                  // vectors per (dest,row) in the receive buffer
    const long long total = static_cast<long long>(world) * n_local * nw * hl_n * hd_vec;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % hd_vec);
        long long t = i / hd_vec;
        const int hl = static_cast<int>(t % hl_n); t /= hl_n;
        const int which = which0 + static_cast<int>(t % nw); t /= nw;
        const long long row = t % n_local;
        const int dest = static_cast<int>(t / n_local);
        const long long src = row * ld_vec + (static_cast<long long>(which) * heads + dest * hl_n + hl) * hd_vec + c;
        const long long dst = (static_cast<long long>(rank) * n_local + row) * per_row + (static_cast<long long>(which) * hl_n + hl) * hd_vec + c;
        recv.p[dest][dst] = qkv[src];
    }
}

// out[row][src*Dl + c] = recv[src][row][c]
__global__ void __launch_bounds__(256)
ulysses_unpack_kernel(const uint4* __restrict__ recv, uint4* __restrict__ out, long long ldo_vec, long long n_local,
                      int dl_vec, int world) {
    const long long total = static_cast<long long>(world) * n_local * dl_vec;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % dl_vec);
        long long t = i / dl_vec;
        const long long row = t % n_local;
        const int src = static_cast<int>(t / n_local);
        out[row * ldo_vec + static_cast<long long>(src) * dl_vec + c] = recv[i];
    }
}

inline unsigned stream_grid(long long work_items, int block) {
    long long g = (work_items + block - 1) / block;
    const long long cap = static_cast<long long>(sm_count()) * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<unsigned>(g);
}

}  // namespace ew
}  // namespace wvd

using namespace wvd;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" __attribute__((visibility("default"))) int wvd_ln_modulate(const void* x, int64_t ldx, const void* shift, const void* scale, const void* weight,
                               const void* bias, void* out, int64_t ldo, int64_t n_tokens, int dim, float eps,
                               int dtype, wvd_stream_t stream) {
    WVD_REQUIRE(x && out, "wvd_ln_modulate: null pointer");
    WVD_REQUIRE((shift == nullptr) == (scale == nullptr), "wvd_ln_modulate: shift and scale go together");
    WVD_REQUIRE((weight == nullptr) == (bias == nullptr), "wvd_ln_modulate: weight and bias go together");
    WVD_REQUIRE(dtype == WVD_BF16 || dtype == WVD_F32, "wvd_ln_modulate: bad dtype %d", dtype);
    const int ve = dtype == WVD_BF16 ? 8 : 4;
    WVD_REQUIRE(dim > 0 && dim % ve == 0 && ldx % ve == 0 && ldo % ve == 0 && ldx >= dim && ldo >= dim,
                "wvd_ln_modulate: dim/ld must be multiples of %d (dim=%d)", ve, dim);
    WVD_REQUIRE(aligned16(x) && aligned16(out) && aligned16(shift) && aligned16(scale) && aligned16(weight) && aligned16(bias),
                "wvd_ln_modulate: pointers must be 16-byte aligned");
    if (n_tokens == 0) return WVD_OK;
    WVD_REQUIRE(n_tokens > 0, "wvd_ln_modulate: negative n_tokens");
    const int nvec = dim / ve;
    const int need = (nvec + 31) / 32;
    cudaStream_t s = (cudaStream_t)stream;
#define WVD_DISPATCH(T)                                                                                             \
    if (need <= 2) return ew::launch_ln<T, 2>(x, ldx, shift, scale, weight, bias, out, ldo, n_tokens, dim, eps, s);     \
    if (need <= 6) return ew::launch_ln<T, 6>(x, ldx, shift, scale, weight, bias, out, ldo, n_tokens, dim, eps, s);     \
    if (need <= 12) return ew::launch_ln<T, 12>(x, ldx, shift, scale, weight, bias, out, ldo, n_tokens, dim, eps, s);   \
    if (need <= 20) return ew::launch_ln<T, 20>(x, ldx, shift, scale, weight, bias, out, ldo, n_tokens, dim, eps, s);   \
    if (need <= 40) return ew::launch_ln<T, 40>(x, ldx, shift, scale, weight, bias, out, ldo, n_tokens, dim, eps, s);
    if (dtype == WVD_BF16) { WVD_DISPATCH(__nv_bfloat16) } else { WVD_DISPATCH(float) }
#undef WVD_DISPATCH
    return set_error(WVD_ERR_UNSUPPORTED, "wvd_ln_modulate: dim %d too large", dim);
}

extern "C" __attribute__((visibility("default"))) int wvd_qk_rmsnorm_rope(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* wq,
                                   const void* wk, void* q_out, int64_t ldqo, void* k_out, int64_t ldko,
                                   int64_t n_tokens, int dim, int head_dim, float eps, const void* rope_cs,
                                   const int32_t* frame_ids, int grid_f, int grid_h, int grid_w, int64_t token_offset,
                                   int dtype, wvd_stream_t stream) {
    WVD_REQUIRE(q && wq && q_out, "wvd_qk_rmsnorm_rope: null pointer");
    WVD_REQUIRE(k == nullptr || (wk && k_out), "wvd_qk_rmsnorm_rope: k needs wk and k_out");
    WVD_REQUIRE(dtype == WVD_BF16 || dtype == WVD_F32, "wvd_qk_rmsnorm_rope: bad dtype %d", dtype);
    WVD_REQUIRE(head_dim == 128, "wvd_qk_rmsnorm_rope: head_dim must be 128 (got %d)", head_dim);
    const int ve = dtype == WVD_BF16 ? 8 : 4;
    WVD_REQUIRE(dim > 0 && dim % head_dim == 0 && ldq % ve == 0 && ldqo % ve == 0 && ldq >= dim && ldqo >= dim,
                "wvd_qk_rmsnorm_rope: bad dim/ld");
    if (k) WVD_REQUIRE(ldk % ve == 0 && ldko % ve == 0 && ldk >= dim && ldko >= dim, "wvd_qk_rmsnorm_rope: bad k ld");
    WVD_REQUIRE(aligned16(q) && aligned16(k) && aligned16(wq) && aligned16(wk) && aligned16(q_out) && aligned16(k_out),
                "wvd_qk_rmsnorm_rope: pointers must be 16-byte aligned");
    if (n_tokens == 0) return WVD_OK;
    WVD_REQUIRE(n_tokens > 0, "wvd_qk_rmsnorm_rope: negative n_tokens");
    if (rope_cs && grid_f == 0) {
        // per-token cos/sin table (n_tokens, 64) float2: the reference's complex `freqs` argument of DiTBlock.forward
        WVD_REQUIRE(grid_h == 0 && grid_w == 0 && frame_ids == nullptr && token_offset == 0,
                    "wvd_qk_rmsnorm_rope: per-token rope table (grid 0x0x0) takes no grid / frame_ids / token_offset");
        grid_h = grid_w = 1;
    } else if (rope_cs) {
        WVD_REQUIRE(grid_f > 0 && grid_h > 0 && grid_w > 0 && grid_h <= 1024 && grid_w <= 1024,
                    "wvd_qk_rmsnorm_rope: bad token grid %dx%dx%d", grid_f, grid_h, grid_w);
        // rows in [grid, token_offset + n_tokens) are tolerated only as the padding of the LAST shard: fewer than one
        // shard's worth of them (n_tokens), the reference's pad-to-ceil(N/P) (wan_video_new.py:1414-1416)
        WVD_REQUIRE(token_offset >= 0 && token_offset <= (int64_t)grid_f * grid_h * grid_w,
                    "wvd_qk_rmsnorm_rope: token_offset %lld is outside the %dx%dx%d grid", (long long)token_offset,
                    grid_f, grid_h, grid_w);
        WVD_REQUIRE(frame_ids != nullptr || grid_f <= 1024, "wvd_qk_rmsnorm_rope: more than 1024 frames");
    }
    const int need = (dim / ve + 31) / 32;
    cudaStream_t s = (cudaStream_t)stream;
#define WVD_DISPATCH(T)                                                                                                \
    if (need <= 2) return ew::launch_rms<T, 2>(q, ldq, k, ldk, wq, wk, q_out, ldqo, k_out, ldko, n_tokens, dim, eps, rope_cs, frame_ids, grid_f, grid_h, grid_w, token_offset, s);   \
    if (need <= 6) return ew::launch_rms<T, 6>(q, ldq, k, ldk, wq, wk, q_out, ldqo, k_out, ldko, n_tokens, dim, eps, rope_cs, frame_ids, grid_f, grid_h, grid_w, token_offset, s);   \
    if (need <= 12) return ew::launch_rms<T, 12>(q, ldq, k, ldk, wq, wk, q_out, ldqo, k_out, ldko, n_tokens, dim, eps, rope_cs, frame_ids, grid_f, grid_h, grid_w, token_offset, s); \
    if (need <= 20) return ew::launch_rms<T, 20>(q, ldq, k, ldk, wq, wk, q_out, ldqo, k_out, ldko, n_tokens, dim, eps, rope_cs, frame_ids, grid_f, grid_h, grid_w, token_offset, s); \
    if (need <= 40) return ew::launch_rms<T, 40>(q, ldq, k, ldk, wq, wk, q_out, ldqo, k_out, ldko, n_tokens, dim, eps, rope_cs, frame_ids, grid_f, grid_h, grid_w, token_offset, s);
    if (dtype == WVD_BF16) { WVD_DISPATCH(__nv_bfloat16) } else { WVD_DISPATCH(float) }
#undef WVD_DISPATCH
    return set_error(WVD_ERR_UNSUPPORTED, "wvd_qk_rmsnorm_rope: dim %d too large", dim);
}

extern "C" __attribute__((visibility("default"))) int wvd_scale_add(const void* x, const void* y, float scale, void* out, int64_t n_elems, int dtype,
                             wvd_stream_t stream) {
    WVD_REQUIRE(dtype == WVD_BF16 || dtype == WVD_F32, "wvd_scale_add: bad dtype %d", dtype);
    if (n_elems == 0) return WVD_OK;
    WVD_REQUIRE(x && y && out && n_elems > 0, "wvd_scale_add: null pointer");
    const int ve = dtype == WVD_BF16 ? 8 : 4;
    WVD_REQUIRE(n_elems % ve == 0, "wvd_scale_add: n_elems must be a multiple of %d", ve);
    WVD_REQUIRE(aligned16(x) && aligned16(y) && aligned16(out), "wvd_scale_add: pointers must be 16-byte aligned");
    const long long nvec = n_elems / ve;
    const unsigned grid = ew::stream_grid(nvec, 256);
    if (dtype == WVD_BF16)
        ew::scale_add_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)y, scale, (__nv_bfloat16*)out, nvec);
    else
        ew::scale_add_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (const float*)y, scale, (float*)out, nvec);
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}

extern "C" __attribute__((visibility("default"))) int wvd_cfg_euler_step(const void* x, const void* v_posi, const void* v_nega, float cfg_scale,
                                  float dsigma, void* out, int64_t n_elems, int dtype, wvd_stream_t stream) {
    WVD_REQUIRE(dtype == WVD_BF16 || dtype == WVD_F32, "wvd_cfg_euler_step: bad dtype %d", dtype);
    if (n_elems == 0) return WVD_OK;
    WVD_REQUIRE(x && v_posi && out && n_elems > 0, "wvd_cfg_euler_step: null pointer");
    const int ve = dtype == WVD_BF16 ? 8 : 4;
    WVD_REQUIRE(n_elems % ve == 0, "wvd_cfg_euler_step: n_elems must be a multiple of %d", ve);
    WVD_REQUIRE(aligned16(x) && aligned16(v_posi) && aligned16(v_nega) && aligned16(out), "wvd_cfg_euler_step: pointers must be 16-byte aligned");
    const long long nvec = n_elems / ve;
    const unsigned grid = ew::stream_grid(nvec, 256);
    if (dtype == WVD_BF16)
        ew::cfg_euler_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)v_posi, (const __nv_bfloat16*)v_nega, cfg_scale, dsigma, (__nv_bfloat16*)out, nvec);
    else
        ew::cfg_euler_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (const float*)v_posi, (const float*)v_nega, cfg_scale, dsigma, (float*)out, nvec);
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}

extern "C" __attribute__((visibility("default"))) int wvd_editor_step(
    const void* z_main, const void* z_edit, const void* vp_main, const void* vp_edit, const void* vn_main, const void* vn_edit,
    int64_t v_main_bc_stride, int64_t v_edit_bc_stride, const int* frame_to_key, const int* key_idx, int bc, int t_frames,
    int k_frames, int64_t hw, float cfg_scale, float dt, float alpha, float beta, float dsigma, int euler, void* out_main,
    void* out_edit, int dtype, wvd_stream_t stream) {
    WVD_REQUIRE(dtype == WVD_BF16 || dtype == WVD_F32, "wvd_editor_step: bad dtype %d", dtype);
    WVD_REQUIRE(z_main && z_edit && vp_main && vp_edit && frame_to_key && key_idx && out_main && out_edit, "wvd_editor_step: null pointer");
    WVD_REQUIRE((vn_main == nullptr) == (vn_edit == nullptr), "wvd_editor_step: the negative branch needs both parts");
    WVD_REQUIRE(bc > 0 && t_frames > 0 && k_frames > 0 && k_frames <= t_frames && hw > 0, "wvd_editor_step: bad sizes");
    const int ve = dtype == WVD_BF16 ? 8 : 4;
    WVD_REQUIRE(hw % ve == 0 && v_main_bc_stride % ve == 0 && v_edit_bc_stride % ve == 0,
                "wvd_editor_step: H*W and the velocity strides must be multiples of %d", ve);
    WVD_REQUIRE(v_main_bc_stride >= (int64_t)t_frames * hw && v_edit_bc_stride >= (int64_t)k_frames * hw, "wvd_editor_step: bad velocity strides");
    WVD_REQUIRE(aligned16(z_main) && aligned16(z_edit) && aligned16(vp_main) && aligned16(vp_edit) && aligned16(vn_main) &&
                aligned16(vn_edit) && aligned16(out_main) && aligned16(out_edit), "wvd_editor_step: pointers must be 16-byte aligned");
    const long long total = (long long)bc * (t_frames + k_frames) * (hw / ve);
    const unsigned grid = ew::stream_grid(total, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == WVD_BF16)
        ew::editor_step_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
            (const __nv_bfloat16*)z_main, (const __nv_bfloat16*)z_edit, (const __nv_bfloat16*)vp_main, (const __nv_bfloat16*)vp_edit,
            (const __nv_bfloat16*)vn_main, (const __nv_bfloat16*)vn_edit, v_main_bc_stride, v_edit_bc_stride, frame_to_key, key_idx, bc,
            t_frames, k_frames, hw, cfg_scale, dt, alpha, beta, dsigma, euler, (__nv_bfloat16*)out_main, (__nv_bfloat16*)out_edit);
    else
        ew::editor_step_kernel<float><<<grid, 256, 0, st>>>(
            (const float*)z_main, (const float*)z_edit, (const float*)vp_main, (const float*)vp_edit, (const float*)vn_main,
            (const float*)vn_edit, v_main_bc_stride, v_edit_bc_stride, frame_to_key, key_idx, bc, t_frames, k_frames, hw, cfg_scale, dt,
            alpha, beta, dsigma, euler, (float*)out_main, (float*)out_edit);
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}

extern "C" __attribute__((visibility("default"))) int wvd_tile_blend(void* values, void* weight, const void* tile, int planes, int H, int W,
                                                                     int th, int tw, int y0, int x0, int bounds, int border_h,
                                                                     int border_w, int dtype, wvd_stream_t stream) {
    WVD_REQUIRE(dtype == WVD_BF16 || dtype == WVD_F32, "wvd_tile_blend: bad dtype %d", dtype);
    WVD_REQUIRE(values && weight && tile, "wvd_tile_blend: null pointer");
    WVD_REQUIRE(planes > 0 && H > 0 && W > 0 && th > 0 && tw > 0 && y0 >= 0 && x0 >= 0 && y0 + th <= H && x0 + tw <= W,
                "wvd_tile_blend: the tile (%d x %d at %d, %d) does not fit the %d x %d plane", th, tw, y0, x0, H, W);
    WVD_REQUIRE(bounds >= 0 && bounds < 16 && border_h >= 0 && border_w >= 0, "wvd_tile_blend: bad bounds / border");
    WVD_REQUIRE(((bounds & 3) == 3 || border_h > 0) && ((bounds & 12) == 12 || border_w > 0),
                "wvd_tile_blend: a side that is not a volume boundary needs a positive border width");
    const int ve = dtype == WVD_BF16 ? 8 : 4;
    WVD_REQUIRE((long long)planes * th * tw < (1ll << 31), "wvd_tile_blend: tile too large (planes * th * tw must be < 2^31)");
    const bool vec = tw % ve == 0 && x0 % ve == 0 && W % ve == 0 && aligned16(values) && aligned16(weight) && aligned16(tile);
    const long long total = (long long)planes * th * (vec ? tw / ve : tw);
    const unsigned grid = ew::stream_grid(total, 256);
    cudaStream_t st = (cudaStream_t)stream;
#define WVD_TB(T, V) ew::tile_blend_kernel<T, V><<<grid, 256, 0, st>>>((T*)values, (T*)weight, (const T*)tile, planes, H, W, th, tw, y0, x0, bounds, border_h, border_w)
    if (dtype == WVD_BF16) { if (vec) WVD_TB(__nv_bfloat16, 8); else WVD_TB(__nv_bfloat16, 1); }
    else { if (vec) WVD_TB(float, 4); else WVD_TB(float, 1); }
#undef WVD_TB
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}

extern "C" __attribute__((visibility("default"))) int wvd_tile_finalize(void* values, const void* weight, int planes, int64_t hw, int clamp,
                                                                        float lo, float hi, int dtype, wvd_stream_t stream) {
    WVD_REQUIRE(dtype == WVD_BF16 || dtype == WVD_F32, "wvd_tile_finalize: bad dtype %d", dtype);
    WVD_REQUIRE(values && weight && planes > 0 && hw > 0, "wvd_tile_finalize: bad arguments");
    const int ve = dtype == WVD_BF16 ? 8 : 4;
    const bool vec = hw % ve == 0 && aligned16(values) && aligned16(weight);
    const long long total = (long long)planes * (vec ? hw / ve : hw);
    const unsigned grid = ew::stream_grid(total, 256);
    cudaStream_t st = (cudaStream_t)stream;
#define WVD_TF(T, V) ew::tile_finalize_kernel<T, V><<<grid, 256, 0, st>>>((T*)values, (const T*)weight, planes, hw, clamp, lo, hi)
    if (dtype == WVD_BF16) { if (vec) WVD_TF(__nv_bfloat16, 8); else WVD_TF(__nv_bfloat16, 1); }
    else { if (vec) WVD_TF(float, 4); else WVD_TF(float, 1); }
#undef WVD_TF
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}

extern "C" __attribute__((visibility("default"))) int wvd_gate_residual(const void* x, const void* gate, const void* y, void* out, int64_t n_tokens, int dim,
                                 int dtype, wvd_stream_t stream) {
    WVD_REQUIRE(dtype == WVD_BF16 || dtype == WVD_F32, "wvd_gate_residual: bad dtype %d", dtype);
    if (n_tokens == 0) return WVD_OK;
    WVD_REQUIRE(x && gate && y && out && n_tokens > 0, "wvd_gate_residual: null pointer");
    const int ve = dtype == WVD_BF16 ? 8 : 4;
    WVD_REQUIRE(dim > 0 && dim % ve == 0, "wvd_gate_residual: dim must be a multiple of %d", ve);
    WVD_REQUIRE(aligned16(x) && aligned16(y) && aligned16(out) && aligned16(gate), "wvd_gate_residual: alignment");
    const long long nvec = n_tokens * (dim / ve);
    const unsigned grid = ew::stream_grid(nvec, 256);
    if (dtype == WVD_BF16)
        ew::gate_residual_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)gate, (const __nv_bfloat16*)y, (__nv_bfloat16*)out, n_tokens, dim);
    else
        ew::gate_residual_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (const float*)gate, (const float*)y, (float*)out, n_tokens, dim);
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}

extern "C" __attribute__((visibility("default"))) int wvd_ulysses_pack_qkv(const void* qkv, int64_t ld, void* send, int64_t n_local, int heads, int head_dim,
                                    int world, wvd_stream_t stream) {
    WVD_REQUIRE(world >= 1 && heads > 0 && heads % world == 0, "wvd_ulysses_pack_qkv: heads (%d) must divide by world (%d)", heads, world);
    WVD_REQUIRE(head_dim % 8 == 0 && ld % 8 == 0 && ld >= 3ll * heads * head_dim, "wvd_ulysses_pack_qkv: bad head_dim/ld");
    if (n_local == 0) return WVD_OK;
    WVD_REQUIRE(qkv && send && n_local > 0 && aligned16(qkv) && aligned16(send), "wvd_ulysses_pack_qkv: bad pointers");
    const long long total = (long long)n_local * 3 * heads * (head_dim / 8);
    ew::ulysses_pack_kernel<<<ew::stream_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)qkv, ld / 8, (uint4*)send, n_local, heads, head_dim / 8, world);
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}

static int scatter_impl(const void* qkv, int64_t ld, void* const* recv_ptrs, int64_t n_local, int heads, int head_dim, int world,
                        int rank, int which0, wvd_stream_t stream) {
    WVD_REQUIRE(world >= 1 && world <= WVD_MAX_PEERS && rank >= 0 && rank < world, "wvd_ulysses_scatter_qkv: bad world/rank %d/%d", world, rank);
    WVD_REQUIRE(heads > 0 && heads % world == 0, "wvd_ulysses_scatter_qkv: heads (%d) must divide by world (%d)", heads, world);
    WVD_REQUIRE(head_dim % 8 == 0 && ld % 8 == 0 && ld >= 3ll * heads * head_dim, "wvd_ulysses_scatter_qkv: bad head_dim/ld");
    if (n_local == 0) return WVD_OK;
    WVD_REQUIRE(qkv && recv_ptrs && n_local > 0 && aligned16(qkv), "wvd_ulysses_scatter_qkv: bad pointers");
    ew::PeerPtrs pp;
    for (int r = 0; r < WVD_MAX_PEERS; ++r) {
        pp.p[r] = r < world ? (uint4*)recv_ptrs[r] : nullptr;
        WVD_REQUIRE(r >= world || (pp.p[r] && aligned16(pp.p[r])), "wvd_ulysses_scatter_qkv: bad receive pointer of rank %d", r);
    }
    const long long total = (long long)n_local * (3 - which0) * heads * (head_dim / 8);
    ew::ulysses_scatter_kernel<<<ew::stream_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)qkv, ld / 8, pp, n_local, heads, head_dim / 8, world, rank, which0);
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}

extern "C" __attribute__((visibility("default"))) int wvd_ulysses_scatter_qkv(const void* qkv, int64_t ld, void* const* recv_ptrs, int64_t n_local, int heads,
                                       int head_dim, int world, int rank, wvd_stream_t stream) {
    return scatter_impl(qkv, ld, recv_ptrs, n_local, heads, head_dim, world, rank, 0, stream);
}

extern "C" __attribute__((visibility("default"))) int wvd_ulysses_scatter_v(const void* qkv, int64_t ld, void* const* recv_ptrs, int64_t n_local, int heads,
                                     int head_dim, int world, int rank, wvd_stream_t stream) {
    return scatter_impl(qkv, ld, recv_ptrs, n_local, heads, head_dim, world, rank, 2, stream);
}

extern "C" __attribute__((visibility("default"))) int wvd_qk_rmsnorm_rope_scatter(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* wq,
                                           const void* wk, void* const* recv_ptrs, int64_t n_local, int dim, int head_dim,
                                           float eps, const void* rope_cs, const int32_t* frame_ids, int grid_f, int grid_h,
                                           int grid_w, int64_t token_offset, int world, int rank, wvd_stream_t stream) {
    WVD_REQUIRE(q && k && wq && wk && recv_ptrs && rope_cs, "wvd_qk_rmsnorm_rope_scatter: null pointer");
    WVD_REQUIRE(head_dim == 128 && dim > 0 && dim % 128 == 0, "wvd_qk_rmsnorm_rope_scatter: head_dim must be 128 and divide dim");
    WVD_REQUIRE(world >= 1 && world <= WVD_MAX_PEERS && rank >= 0 && rank < world, "wvd_qk_rmsnorm_rope_scatter: bad world/rank %d/%d", world, rank);
    const int heads = dim / 128;
    WVD_REQUIRE(heads % world == 0, "wvd_qk_rmsnorm_rope_scatter: heads (%d) must divide by world (%d)", heads, world);
    WVD_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldq >= dim && ldk >= dim, "wvd_qk_rmsnorm_rope_scatter: bad ld");
    WVD_REQUIRE(aligned16(q) && aligned16(k) && aligned16(wq) && aligned16(wk), "wvd_qk_rmsnorm_rope_scatter: pointers must be 16-byte aligned");
    if (n_local == 0) return WVD_OK;
    WVD_REQUIRE(n_local > 0, "wvd_qk_rmsnorm_rope_scatter: negative n_local");
    WVD_REQUIRE(grid_f > 0 && grid_h > 0 && grid_w > 0 && grid_h <= 1024 && grid_w <= 1024 && (frame_ids != nullptr || grid_f <= 1024),
                "wvd_qk_rmsnorm_rope_scatter: bad token grid %dx%dx%d", grid_f, grid_h, grid_w);
    WVD_REQUIRE(token_offset >= 0 && token_offset <= (int64_t)grid_f * grid_h * grid_w, "wvd_qk_rmsnorm_rope_scatter: token_offset outside the grid");
    ew::PeerPtrs pp;
    for (int r = 0; r < WVD_MAX_PEERS; ++r) {
        pp.p[r] = r < world ? (uint4*)recv_ptrs[r] : nullptr;
        WVD_REQUIRE(r >= world || (pp.p[r] && aligned16(pp.p[r])), "wvd_qk_rmsnorm_rope_scatter: bad receive pointer of rank %d", r);
    }
    const int need = (dim / 8 + 31) / 32;
    const dim3 grid(static_cast<unsigned>((n_local + ew::WARPS_PER_BLOCK - 1) / ew::WARPS_PER_BLOCK), 2);
    const dim3 block(ew::WARPS_PER_BLOCK * 32);
    cudaStream_t s = (cudaStream_t)stream;
#define WVD_LAUNCH(MV)                                                                                                  \
    ew::qk_rmsnorm_rope_scatter_kernel<MV><<<grid, block, 0, s>>>((const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, \
        (const __nv_bfloat16*)wq, (const __nv_bfloat16*)wk, pp, n_local, dim, eps, (const float2*)rope_cs, frame_ids, grid_f,  \
        grid_h, grid_w, token_offset, heads / world, rank)
    if (need <= 2) WVD_LAUNCH(2);
    else if (need <= 6) WVD_LAUNCH(6);
    else if (need <= 12) WVD_LAUNCH(12);
    else if (need <= 20) WVD_LAUNCH(20);
    else if (need <= 40) WVD_LAUNCH(40);
    else return set_error(WVD_ERR_UNSUPPORTED, "wvd_qk_rmsnorm_rope_scatter: dim %d too large", dim);
#undef WVD_LAUNCH
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}

extern "C" __attribute__((visibility("default"))) int wvd_ulysses_unpack_out(const void* recv, void* out, int64_t ldo, int64_t n_local, int heads,
                                      int head_dim, int world, wvd_stream_t stream) {
    WVD_REQUIRE(world >= 1 && heads > 0 && heads % world == 0, "wvd_ulysses_unpack_out: heads (%d) must divide by world (%d)", heads, world);
    WVD_REQUIRE(head_dim % 8 == 0 && ldo % 8 == 0 && ldo >= (int64_t)heads * head_dim, "wvd_ulysses_unpack_out: bad head_dim/ld");
    if (n_local == 0) return WVD_OK;
    WVD_REQUIRE(recv && out && n_local > 0 && aligned16(recv) && aligned16(out), "wvd_ulysses_unpack_out: bad pointers");
    const int dl_vec = (heads / world) * head_dim / 8;
    const long long total = (long long)world * n_local * dl_vec;
    ew::ulysses_unpack_kernel<<<ew::stream_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)recv, (uint4*)out, ldo / 8, n_local, dl_vec, world);
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}
