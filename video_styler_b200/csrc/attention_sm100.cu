// K8/K9: flash-style attention forward for head_dim 128 on tcgen05 / TMEM (sm_100a).
//
//   out[s, h] = softmax(q_h k_h^T * scale) v_h          non-causal, no mask, no dropout
//   (replaces flash_attention(), diffsynth/models/wan_video_dit.py:28-61; self: Sk = Sq ~ 30k-76k, cross: Sk = 512)
//
// One CTA = one head x 256 query rows (two 128-row Q tiles, ping-pong), 384 threads:
//   warps 0-3 / 4-7     softmax warpgroup for Q tile 0 / 1: one query row per thread; S row TMEM->registers,
//                       running max with lazy rescale (only when the max grows by > 2^8), exp2 (MUFU, packed fp32x2
//                       scale/sum), P (bf16) written back to TMEM over the S columns, row sum in fp32; final O / l -> global
//   warp 8 (1 thread)   TMA producer: Q tiles once, then K_j, V_j (128x128 bf16, 128-byte swizzle) through a 4-slot ring
//   warp 9 (1 thread)   MMA issuer  : S_i = Q_i K_j^T   (SS, K-major A and B, 128x128x16 x8) -> TMEM
//                                     O_i += P_i V_j    (TS: P from TMEM, V MN-major from smem, 128x128x16 x8) -> TMEM
//   warp 10             TMEM allocator (512 columns)
// While warpgroup i does softmax on S_i(j), the tensor core runs PV / QK^T of the other tile.
// TMEM map (512 columns): S0 [0,128) | S1 [128,256) | O0 [256,384) | O1 [384,512); P_i aliases S_i[0,64).
// KV tail (Sk % 128 != 0): TMA zero-fills, the softmax masks the tail columns to -inf.  Q tail rows are not stored.
// Measured structure of one KV step (ncu + in-kernel cycle counters, profiles/): tensor pipe 2048 cycles nominal,
// MUFU.EX2 2048 cycles (16/clk/SM) -- the two pipes are co-critical and coupled through the S->P->S dependency
// chain of each tile, which is what bounds the kernel at ~56 % tensor-pipe activity.
#include <math.h>
#include <stdlib.h>

#include "host_utils.h"
#include "ptx.cuh"

namespace wvd {
namespace attn {

constexpr int BQ = 128;               // rows per Q tile
constexpr int QT = 2;                 // Q tiles per CTA
constexpr int BKV = 128;              // keys per KV tile
constexpr int HD = 128;               // head dim
constexpr int TILE_BYTES = 128 * 128 * 2;     // 32 KB
constexpr int HALF_BYTES = TILE_BYTES / 2;    // one 64-column TMA box
constexpr int SLOTS = 4;              // K/V ring slots
constexpr int NUM_THREADS = 384;
constexpr int SMEM_BYTES = QT * TILE_BYTES + SLOTS * TILE_BYTES + 1024 + 256;
constexpr uint32_t IDESC_QK = make_idesc_bf16(128, 128, 0, 0);   // A = Q (K-major), B = K (K-major)
constexpr uint32_t IDESC_PV = make_idesc_bf16(128, 128, 0, 1);   // A = P (TMEM), B = V (MN-major)
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 units
// Warp roles.  The control warps get the HIGHEST warp ids: the SM's issue arbiter favours higher warp ids, and the
// single MMA-issuing thread must never wait behind the eight softmax warps for an issue slot.
#ifndef WVD_ATTN_NCHUNK
#define WVD_ATTN_NCHUNK 1
#endif
constexpr int NCHUNK = WVD_ATTN_NCHUNK;     // P is produced / consumed in NCHUNK chunks of 128/NCHUNK keys (1, 2 or 4)
constexpr int CTRL_WARP0 = 8, TMA_WARP = 8, MMA_WARP = 9, ALLOC_WARP = 10;

struct Params {
    __nv_bfloat16* out;
    long long ldo;
    int sq, sk, n_kv;
    float scale_log2;
    unsigned long long* prof;   // optional device buffer (developer profiling, see wvd_debug_attention_profile)
    int dbg;                    // developer timing experiments (WVD_ATTN_DEBUG): 1 = skip softmax math, 2 = skip QK MMAs, 4 = skip PV MMAs
};

__device__ __forceinline__ uint32_t clk32() {
    uint32_t c;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
    return c;
}

// EMU_OF_4: how many of every 4 consecutive column pairs take exp2 on the FMA pipes (polynomial) instead of MUFU.
// MUFU.EX2 runs at 16/clk/SM: 256 rows x 128 columns per KV step = 2048 cycles, exactly the tensor-pipe time of the
// step; moving part of the exponentials to the FMA pipes takes the XU pipe off the critical path.
template <int EMU_OF_4>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - raw_u32);
    const uint32_t q_smem = smem_base;                                  // QT tiles
    const uint32_t kv_smem = smem_base + QT * TILE_BYTES;               // SLOTS tiles
    const uint32_t bar_base = kv_smem + SLOTS * TILE_BYTES;
    const uint32_t q_full = bar_base;
    auto kv_full = [&](int s) { return bar_base + 8 + s * 8; };
    auto kv_empty = [&](int s) { return bar_base + 40 + s * 8; };
    auto s_full = [&](int i) { return bar_base + 72 + i * 8; };
    auto o_full = [&](int i) { return bar_base + 88 + i * 8; };
    auto p_full = [&](int i, int c) { return bar_base + 104 + (i * NCHUNK + c) * 8; };   // P chunk c of tile i is in TMEM
    const uint32_t tmem_slot = bar_base + 176;
    volatile uint32_t* tmem_slot_gen =
        reinterpret_cast<volatile uint32_t*>(smem_gen + (QT + SLOTS) * TILE_BYTES + 176);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int head = blockIdx.y;
    const int q_row0 = blockIdx.x * (QT * BQ);
    const int n_kv = p.n_kv;

    if (warp == TMA_WARP && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == MMA_WARP && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < SLOTS; ++s) {
            mbar_init(kv_full(s), 1);
            mbar_init(kv_empty(s), 1);
        }
        for (int i = 0; i < QT; ++i) {
            mbar_init(s_full(i), 1);
            for (int c = 0; c < NCHUNK; ++c) mbar_init(p_full(i, c), 4);      // one arrival per softmax warp
            mbar_init(o_full(i), 1);
        }
        fence_barrier_init();
    }
    if (warp == ALLOC_WARP) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp >= CTRL_WARP0) {
        setmaxnreg_dec<72>();
        if (warp == TMA_WARP && lane == 0) {
            // ------------------------------ TMA producer ------------------------------
            mbar_expect_tx(q_full, QT * TILE_BYTES);
#pragma unroll
            for (int i = 0; i < QT; ++i) {
                tma_load_2d(q_smem + i * TILE_BYTES, &tmQ, q_full, head * HD, q_row0 + i * BQ);
                tma_load_2d(q_smem + i * TILE_BYTES + HALF_BYTES, &tmQ, q_full, head * HD + 64, q_row0 + i * BQ);
            }
            for (int t = 0; t < 2 * n_kv; ++t) {
                const int slot = t % SLOTS;
                const uint32_t ph = (t / SLOTS) & 1;
                const int j = t >> 1;
                const CUtensorMap* tm = (t & 1) ? &tmV : &tmK;
                mbar_wait(kv_empty(slot), ph ^ 1, 0x100 + slot);
                mbar_expect_tx(kv_full(slot), TILE_BYTES);
                const uint32_t dst = kv_smem + slot * TILE_BYTES;
                tma_load_2d(dst, tm, kv_full(slot), head * HD, j * BKV);
                tma_load_2d(dst + HALF_BYTES, tm, kv_full(slot), head * HD + 64, j * BKV);
            }
        } else if (warp == MMA_WARP && lane == 0) {
            // ------------------------------ MMA issuer ------------------------------
            // Every mbarrier probe costs ~100 cycles of latency on this single thread, so the schedule is a fixed
            // ping-pong with as few waits as possible (an event-driven poller over both tiles measured 2x slower).
            auto issue_qk = [&](int i, uint32_t k_addr) {
                if (p.dbg & 2) return;
                const uint32_t qa = q_smem + i * TILE_BYTES;
                const uint32_t d = tmem_base + i * 128;
#pragma unroll
                for (int kk = 0; kk < HD / 16; ++kk) {
                    const uint32_t off = (kk >> 2) * HALF_BYTES + (kk & 3) * 32;
                    umma_ss(d, make_smem_desc_sw128(qa + off, 16, 1024), make_smem_desc_sw128(k_addr + off, 16, 1024),
                            IDESC_QK, kk != 0 ? 1u : 0u);
                }
            };
            auto issue_pv = [&](int i, uint32_t v_addr, bool accumulate, uint32_t pph) {
                const uint32_t d = tmem_base + 256 + i * 128;
                const uint32_t pa = tmem_base + i * 128;
#pragma unroll
                for (int c = 0; c < NCHUNK; ++c) {
                    mbar_wait(p_full(i, c), pph, 0x230 + i * 8 + c);
                    tc_fence_after();
                    if (p.dbg & 4) continue;
#pragma unroll
                    for (int kk = c * (8 / NCHUNK); kk < (c + 1) * (8 / NCHUNK); ++kk) {
                        // V tile: kv rows of 128 B (64 d-columns) per box; 16 kv rows = 2048 B; second d-half at +16 KB
                        umma_ts(d, pa + kk * 8, make_smem_desc_sw128(v_addr + kk * 2048, HALF_BYTES, 1024), IDESC_PV,
                                (accumulate || kk != 0) ? 1u : 0u);
                    }
                }
            };
            auto slot_of = [&](int t) { return t % SLOTS; };
            auto phase_of = [&](int t) { return static_cast<uint32_t>((t / SLOTS) & 1); };

            mbar_wait(q_full, 0, 0x200);
            mbar_wait(kv_full(slot_of(0)), phase_of(0), 0x210);
            tc_fence_after();
            issue_qk(0, kv_smem + slot_of(0) * TILE_BYTES);
            tc_commit(s_full(0));
            issue_qk(1, kv_smem + slot_of(0) * TILE_BYTES);
            tc_commit(s_full(1));
            tc_commit(kv_empty(slot_of(0)));
            for (int j = 0; j < n_kv; ++j) {
                const int tv = 2 * j + 1, tk = 2 * j + 2;
                const uint32_t pph = j & 1;
                const bool more = j + 1 < n_kv;
                mbar_wait(kv_full(slot_of(tv)), phase_of(tv), 0x220);
                const uint32_t v_addr = kv_smem + slot_of(tv) * TILE_BYTES;
                const uint32_t k_addr = kv_smem + slot_of(tk) * TILE_BYTES;
                // tile 0
                issue_pv(0, v_addr, j > 0, pph);
                if (more) {
                    mbar_wait(kv_full(slot_of(tk)), phase_of(tk), 0x240);
                    tc_fence_after();
                    issue_qk(0, k_addr);
                    tc_commit(s_full(0));
                } else {
                    tc_commit(o_full(0));
                }
                // tile 1
                issue_pv(1, v_addr, j > 0, pph);
                tc_commit(kv_empty(slot_of(tv)));
                if (more) {
                    issue_qk(1, k_addr);
                    tc_commit(s_full(1));
                    tc_commit(kv_empty(slot_of(tk)));
                } else {
                    tc_commit(o_full(1));
                }
            }
        }
    } else {
        // ------------------------------ softmax warpgroups ------------------------------
        setmaxnreg_inc<216>();
        const int i = warp >> 2;                    // Q tile of this warpgroup (warps 0-3 / 4-7)
        const int quarter = warp & 3;               // TMEM lane quarter accessible to this warp
        const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t s_tmem = tmem_base + i * 128 + lane_sel;
        const uint32_t o_tmem = tmem_base + 256 + i * 128 + lane_sel;
        const int row = q_row0 + i * BQ + quarter * 32 + lane;
        const float sl2 = p.scale_log2;
        const int tail_valid = p.sk - (n_kv - 1) * BKV;     // valid keys in the last KV tile (1..128)
        float m = -INFINITY;   // running max of the raw scores (reference point of the stored exponentials)
        float l = 0.f;

        const bool prof = p.prof != nullptr && blockIdx.x == 1 && blockIdx.y == 0 && lane == 0;
        uint32_t pc_wait = 0, pc_ld = 0, pc_max = 0, pc_exp = 0, pc_st = 0, pt = 0;
        for (int j = 0; j < n_kv; ++j) {
            if (prof) pt = clk32();
            mbar_wait(s_full(i), j & 1, 0x300 + i);
            tc_fence_after();
            if (prof) { const uint32_t t = clk32(); pc_wait += t - pt; pt = t; }
            if (p.dbg & 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { for (int c = 0; c < NCHUNK; ++c) mbar_arrive(p_full(i, c)); }
                continue;
            }
            uint32_t s[128];
            tmem_ld_32x32b_x32(s_tmem + 0, s + 0);
            tmem_ld_32x32b_x32(s_tmem + 32, s + 32);
            tmem_ld_32x32b_x32(s_tmem + 64, s + 64);
            tmem_ld_32x32b_x32(s_tmem + 96, s + 96);
            tc_wait_ld();
            if (prof) { const uint32_t t = clk32(); pc_ld += t - pt; pt = t; }
            if (j == n_kv - 1 && tail_valid < BKV) {
#pragma unroll
                for (int c = 0; c < 128; ++c)
                    if (c >= tail_valid) s[c] = 0xff800000u;   // -inf
            }
            float mx[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) mx[a] = fmaxf(__uint_as_float(s[2 * a]), __uint_as_float(s[2 * a + 1]));
#pragma unroll
            for (int c = 16; c < 128; c += 16) {
#pragma unroll
                for (int a = 0; a < 8; ++a) mx[a] = fmax3(mx[a], __uint_as_float(s[c + 2 * a]), __uint_as_float(s[c + 2 * a + 1]));
            }
            const float m_new = fmaxf(fmax3(fmax3(mx[0], mx[1], mx[2]), fmax3(mx[3], mx[4], mx[5]), fmaxf(mx[6], mx[7])), m);
            if (j == 0) {
                m = m_new;
            } else {
                const bool grow = (m_new - m) * sl2 > RESCALE_THRESHOLD;
                if (__any_sync(0xffffffffu, grow)) {
                    // O_i(j-1) is complete (s_full(j) was committed after PV_i(j-1)); rescale this row
                    const float alpha = fast_exp2((m - m_new) * sl2);
                    l *= alpha;
                    m = m_new;
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        uint32_t o[32];
                        tmem_ld_32x32b_x32(o_tmem + c * 32, o);
                        tc_wait_ld();
#pragma unroll
                        for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                        tmem_st_32x32b_x32(o_tmem + c * 32, o);
                    }
                    tc_wait_st();
                }
            }
            if (prof) { const uint32_t t = clk32(); pc_max += t - pt; pt = t; }
            const float neg_m = -m * sl2;
            const uint64_t sl2_2 = f2_pack(sl2, sl2), negm_2 = f2_pack(neg_m, neg_m);
            uint64_t lsum_a = f2_pack(0.f, 0.f), lsum_b = f2_pack(0.f, 0.f);
#pragma unroll
            for (int ch = 0; ch < NCHUNK; ++ch) {
                constexpr int CW = BKV / NCHUNK;          // keys per chunk
                uint32_t pk[CW / 2];
#pragma unroll
                for (int cc = 0; cc < CW; cc += 2) {
                    const int c = ch * CW + cc;
                    const uint64_t x2 = f2_fma(f2_pack(__uint_as_float(s[c]), __uint_as_float(s[c + 1])), sl2_2, negm_2);
                    float p0, p1;
                    if (((c >> 1) & 3) < EMU_OF_4) {
                        exp2_poly2(x2, p0, p1);
                    } else {
                        float x0, x1;
                        f2_unpack(x2, x0, x1);
                        p0 = fast_exp2(x0);
                        p1 = fast_exp2(x1);
                    }
                    if (c & 2) lsum_b = f2_add(lsum_b, f2_pack(p0, p1));
                    else lsum_a = f2_add(lsum_a, f2_pack(p0, p1));
                    pk[cc >> 1] = pack_bf16x2(p0, p1);
                }
                if (CW == 32) {
                    tmem_st_32x32b_x16(s_tmem + ch * 16, pk);
                } else {
#pragma unroll
                    for (int q4 = 0; q4 < CW / 64; ++q4) tmem_st_32x32b_x32(s_tmem + ch * (CW / 2) + q4 * 32, pk + q4 * 32);
                }
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full(i, ch));
            }
            {
                float a0, a1;
                f2_unpack(f2_add(lsum_a, lsum_b), a0, a1);
                l += a0 + a1;
            }
            if (prof) { const uint32_t t = clk32(); pc_exp += t - pt; pt = t; }
            if (prof) { const uint32_t t = clk32(); pc_st += t - pt; pt = t; }
        }
        if (prof) {
            unsigned long long* o = p.prof + warp * 8;
            o[0] = pc_wait; o[1] = pc_ld; o[2] = pc_max; o[3] = pc_exp; o[4] = pc_st; o[5] = n_kv;
        }

        // ------------------------------ epilogue: O / l -> global ------------------------------
        mbar_wait(o_full(i), 0, 0x310 + i);
        tc_fence_after();
        const float inv_l = 1.0f / l;
        __nv_bfloat16* orow = p.out + static_cast<long long>(row) * p.ldo + head * HD;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(o_tmem + c * 32, o);
            tc_wait_ld();
            if (row < p.sq) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
                    u.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
                    u.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
                    u.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
                    *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = u;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == ALLOC_WARP) tmem_dealloc(tmem_base, 512);
}

unsigned long long* g_prof_buffer = nullptr;
}  // namespace attn

int attn_read_diag(unsigned long long* out) {
    if (cudaMemcpyFromSymbol(out, g_diag, sizeof(unsigned long long) * 8) != cudaSuccess) return -1;
    unsigned long long zero[8] = {0};
    cudaMemcpyToSymbol(g_diag, zero, sizeof(zero));
    return 0;
}

}  // namespace wvd

extern "C" __attribute__((visibility("default"))) int wvd_debug_attention_profile(unsigned long long* device_buf) {
    wvd::attn::g_prof_buffer = device_buf;
    return WVD_OK;
}

extern "C" __attribute__((visibility("default"))) int wvd_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                 void* out, int64_t ldo, int num_heads, int64_t sq, int64_t sk, int head_dim,
                                 float scale, wvd_stream_t stream) {
    using namespace wvd;
    WVD_REQUIRE(q && k && v && out, "wvd_attention_fwd: null pointer");
    WVD_REQUIRE(head_dim == attn::HD, "wvd_attention_fwd: head_dim must be 128 (got %d)", head_dim);
    WVD_REQUIRE(num_heads > 0 && num_heads <= 65535, "wvd_attention_fwd: bad num_heads %d", num_heads);
    WVD_REQUIRE(sq > 0 && sk > 0 && sq < (1ll << 31) && sk < (1ll << 31), "wvd_attention_fwd: bad sequence lengths sq=%lld sk=%lld", (long long)sq, (long long)sk);
    const int64_t width = (int64_t)num_heads * head_dim;
    WVD_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && ldq >= width && ldk >= width && ldv >= width && ldo >= width,
                "wvd_attention_fwd: leading dims must be multiples of 8 and >= heads*128");
    WVD_REQUIRE(((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0) && ((uintptr_t)out % 16 == 0),
                "wvd_attention_fwd: pointers must be 16-byte aligned");
    CUtensorMap tmQ, tmK, tmV;
    int rc = get_tensor_map_bf16(&tmQ, q, (uint64_t)sq, (uint64_t)width, (uint64_t)ldq, attn::BQ);
    if (rc) return rc;
    rc = get_tensor_map_bf16(&tmK, k, (uint64_t)sk, (uint64_t)width, (uint64_t)ldk, attn::BKV);
    if (rc) return rc;
    rc = get_tensor_map_bf16(&tmV, v, (uint64_t)sk, (uint64_t)width, (uint64_t)ldv, attn::BKV);
    if (rc) return rc;
    attn::Params p;
    p.out = (__nv_bfloat16*)out;
    p.ldo = ldo;
    p.sq = (int)sq;
    p.sk = (int)sk;
    p.n_kv = (int)((sk + attn::BKV - 1) / attn::BKV);
    p.scale_log2 = scale * 1.4426950408889634f;
    p.prof = attn::g_prof_buffer;
    {
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("WVD_ATTN_DEBUG"); dbg = e ? atoi(e) : 0; }
        p.dbg = dbg;
    }
    static int emu = -1;
    if (emu < 0) {
        const char* e = getenv("WVD_ATTN_EMU");        // tuning knob: 0..3 of every 4 column pairs on the FMA pipes
        int v = e ? atoi(e) : 0;          // measured on B200: 0 is fastest (the FMA-pipe polynomial costs ~16 cycles/element)
        emu = v < 0 ? 0 : (v > 3 ? 3 : v);
        WVD_CHECK_CUDA(cudaFuncSetAttribute(attn::attention_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::SMEM_BYTES));
        WVD_CHECK_CUDA(cudaFuncSetAttribute(attn::attention_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::SMEM_BYTES));
        WVD_CHECK_CUDA(cudaFuncSetAttribute(attn::attention_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::SMEM_BYTES));
        WVD_CHECK_CUDA(cudaFuncSetAttribute(attn::attention_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::SMEM_BYTES));
    }
    dim3 grid((unsigned)((sq + attn::QT * attn::BQ - 1) / (attn::QT * attn::BQ)), (unsigned)num_heads);
    cudaStream_t st = (cudaStream_t)stream;
    switch (emu) {
        case 0: attn::attention_fwd_kernel<0><<<grid, attn::NUM_THREADS, attn::SMEM_BYTES, st>>>(tmQ, tmK, tmV, p); break;
        case 1: attn::attention_fwd_kernel<1><<<grid, attn::NUM_THREADS, attn::SMEM_BYTES, st>>>(tmQ, tmK, tmV, p); break;
        case 2: attn::attention_fwd_kernel<2><<<grid, attn::NUM_THREADS, attn::SMEM_BYTES, st>>>(tmQ, tmK, tmV, p); break;
        default: attn::attention_fwd_kernel<3><<<grid, attn::NUM_THREADS, attn::SMEM_BYTES, st>>>(tmQ, tmK, tmV, p); break;
    }
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}
