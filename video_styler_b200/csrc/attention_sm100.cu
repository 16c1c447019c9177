// K8/K9: flash-style attention forward for head_dim 128 on tcgen05 / TMEM (sm_100a).
//
//   out[s, h] = softmax(q_h k_h^T * scale) v_h          non-causal, no mask, no dropout
//   (replaces flash_attention(), diffsynth/models/wan_video_dit.py:28-61; self: Sk = Sq ~ 30k-76k, cross: Sk = 512)
//
// One CTA = one head x 256 query rows (two 128-row Q tiles, ping-pong), 576 threads:
//   warps 0-7 / 8-15    softmax warps of Q tile 0 / 1.  TWO threads per query row (warps w and w+4 of a tile share a
//                       TMEM lane quarter and a scheduler): each owns 64 of the 128 score columns of its row.  A lone
//                       warp cannot keep the MUFU pipe busy (in-order issue exposes the MUFU latency: ~70 % of the
//                       16 exp2/clk/SM, tools/microbench/softmax_chunks.cu); with four softmax warps per scheduler
//                       the pipe is fed whenever either tile has scores, and one tile's exponentials overlap the
//                       other tile's PV / QK^T on the tensor core.
//                       exp2 runs against a STALE reference maximum: the row maximum of the tile is exchanged between
//                       the two threads of the row through shared memory and only guards against overflow (the
//                       reference moves when exceeded by 2^8, which also bounds how often O is rescaled).
//                       P (bf16) goes back to TMEM over the S columns in groups of 16 columns and is handed to the
//                       tensor core in two hand-overs so that PV of the first overlaps the exponentials of the
//                       second; row sums in fp32; final O / l -> global.
//   warp 16 (1 thread)  TMA producer: Q tiles once, then K_j, V_j (128x128 bf16, 128-byte swizzle) through a 4-slot ring
//   warp 17 (1 thread)  MMA issuer  : S_i = Q_i K_j^T   (SS, K-major A and B, 128x128x16 x8) -> TMEM
//                                     O_i += P_i V_j    (TS: P from TMEM, V MN-major from smem, 128x128x16 x8) -> TMEM
//                       (the whole warp allocates / frees the 512 TMEM columns)
// TMEM map (512 columns): S0 [0,128) | S1 [128,256) | O0 [256,384) | O1 [384,512); P_i aliases S_i[0,64).
// KV tail (Sk % 128 != 0): TMA zero-fills, the softmax masks the tail columns to -inf.  Q tail rows are not stored.
// Budget of one KV step per SM: tensor pipe 2048 cycles (4 x 128x128x128), MUFU.EX2 2048 cycles (256 x 128 / 16 per clk).
#include <math.h>
#include <stdlib.h>

#include "host_utils.h"
#include "ptx.cuh"
#include "softmax_math.cuh"

namespace wvd {
namespace attn {

constexpr int BQ = 128;               // rows per Q tile
constexpr int BKV = 128;              // keys per KV tile
constexpr int HD = 128;               // head dim
constexpr int NS = BKV / 2;           // score columns per softmax thread
constexpr int TILE_BYTES = 128 * 128 * 2;     // 32 KB
constexpr int HALF_BYTES = TILE_BYTES / 2;    // one 64-column TMA box
// Two shapes of the same kernel (template parameter QT = Q tiles per CTA):
//   QT = 2   the CTA described above: 256 query rows, all 512 TMEM columns, 4 K/V ring slots, 576 threads, one CTA per SM
//   QT = 1   HALF of it -- one 128-row Q tile, 256 TMEM columns (S | O), 2 ring slots (K in one, V in the other), 320
//            threads, 99 KB of shared memory -- so that TWO CTAs are resident per SM.  For the 512-key text
//            cross-attention (4 KV steps per CTA) the prologue (TMEM allocation, barrier init, Q + first K load, first
//            QK^T) and the epilogue are as long as the main loop; with two independent CTAs per SM one CTA's prologue /
//            epilogue runs under the other's main loop, which is what the ping-pong between the two tiles of ONE CTA
//            cannot give (both tiles start and end together).
template <int QT>
struct Cfg {
    static constexpr int SLOTS = QT == 2 ? 4 : 2;             // K/V ring slots
    // Warp roles.  The control warps get the HIGHEST warp ids: the SM's issue arbiter favours higher warp ids, and the
    // single MMA-issuing thread must never wait behind the softmax warps for an issue slot.
    static constexpr int SOFTMAX_WARPS = 8 * QT;
    static constexpr int TMA_WARP = 8 * QT, MMA_WARP = 8 * QT + 1, ALLOC_WARP = 8 * QT + 1;     // the MMA warp also owns the TMEM allocation
    static constexpr int NUM_THREADS = (8 * QT + 2) * 32;
    static constexpr int TMEM_COLS = 256 * QT;                // S_i (128 each) then O_i (128 each)
    static constexpr int O_COL = 128 * QT;
    static constexpr int XCHG_BYTES = 2 * QT * 2 * BQ * 4;    // [step parity][tile][half][row] fp32
    static constexpr int SMEM_BYTES = QT * TILE_BYTES + SLOTS * TILE_BYTES + 256 + XCHG_BYTES + 1024;
    static constexpr int MIN_CTAS = QT == 2 ? 1 : 2;
};
// Registers: 576 threads x 112 = 64,512 of the SM's 65,536 (the allocation unit is 16 per thread; a 19th warp would
// cap every thread at 96).  setmaxnreg.inc can only take registers that other warps of the CTA have released, and two
// control warps cannot release enough to matter for 512 softmax threads, so the softmax code is written to fit the
// launch allocation (64 score columns + 32 packed P columns per thread).
constexpr int BAR_BYTES = 256;
constexpr uint32_t IDESC_QK = make_idesc_bf16(128, 128, 0, 0);   // A = Q (K-major), B = K (K-major)
constexpr uint32_t IDESC_PV = make_idesc_bf16(128, 128, 0, 1);   // A = P (TMEM), B = V (MN-major)
#ifndef WVD_ATTN_HO0_GROUPS
#define WVD_ATTN_HO0_GROUPS 2
#endif
#ifndef WVD_ATTN_RELEASE_GROUP
#define WVD_ATTN_RELEASE_GROUP 3
#endif
#ifndef WVD_ATTN_KERNEL_DEFAULT
#define WVD_ATTN_KERNEL_DEFAULT 0        // 0 = by key length, 1 = the two-tile kernel of this file, 2 = the CTA-pair kernel (attention_pair_sm100.cu)
#endif
#ifndef WVD_ATTN_TURNS
#define WVD_ATTN_TURNS 0
#endif
constexpr bool TURNS = WVD_ATTN_TURNS != 0;             // make the two tiles take strict turns on the MUFU pipe (measured 3 % slower than free-running, tools/attn_ab.py)
constexpr int GC = 16;                                  // columns per exp2 / store group (4 groups per thread)
constexpr int HO0_GROUPS = WVD_ATTN_HO0_GROUPS;         // groups in the first hand-over of P (the rest form the second)
constexpr int RELEASE_GROUP = WVD_ATTN_RELEASE_GROUP;   // the other tile's turn starts once this group's exponentials are issued
constexpr float REF_MARGIN = 8.0f;    // log2 units: the softmax reference point moves when the row maximum exceeds it by 2^8

struct Params {
    __nv_bfloat16* out;
    long long ldo;
    // Ulysses return trip fused into the epilogue: query row `row` belongs to rank row / rows_per_peer and is stored
    // straight into that rank's buffer (peer pointer, NVLink) at local row row % rows_per_peer; 0 = single output.
    __nv_bfloat16* out_peer[WVD_MAX_PEERS];
    int rows_per_peer;
    int sq, sk, n_kv;
    float scale_log2;
    unsigned long long* prof;   // optional device buffer (developer profiling, -DWVD_ATTN_PROF builds only)
};

__device__ __forceinline__ uint32_t clk32() {
    uint32_t c;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
    return c;
}
#ifdef WVD_ATTN_PROF
#define PROF_LAP(acc) do { if (prof) { const uint32_t t_ = clk32(); (acc) += t_ - pt; pt = t_; } } while (0)
#else
#define PROF_LAP(acc) do { } while (0)
#endif

template <int EMU_OF_4, int QT>
__global__ void __launch_bounds__(Cfg<QT>::NUM_THREADS, Cfg<QT>::MIN_CTAS)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const Params p) {
    constexpr int SLOTS = Cfg<QT>::SLOTS, SOFTMAX_WARPS = Cfg<QT>::SOFTMAX_WARPS, TMA_WARP = Cfg<QT>::TMA_WARP;
    constexpr int MMA_WARP = Cfg<QT>::MMA_WARP, ALLOC_WARP = Cfg<QT>::ALLOC_WARP, O_COL = Cfg<QT>::O_COL;
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
    auto p_full = [&](int i, int c) { return bar_base + 104 + (i * 2 + c) * 8; };   // hand-over c of tile i is in TMEM
    const uint32_t turn_word = bar_base + 160;                                      // always zero (see the softmax loop)
    const uint32_t tmem_slot = bar_base + 176;
    const uint32_t xchg = bar_base + BAR_BYTES;
    volatile uint32_t* tmem_slot_gen =
        reinterpret_cast<volatile uint32_t*>(smem_gen + (QT + SLOTS) * TILE_BYTES + 176);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int head = blockIdx.y;
    const int q_row0 = blockIdx.x * (QT * BQ);
    const int n_kv = p.n_kv;
#ifdef WVD_ATTN_PROF
    // CTA timeline (cycles since entry) of a few CTAs of head 1 (a later wave than head 0): after setup, first S seen,
    // main loop done, O complete, stores done -> p.prof[128 + 8 * blockIdx.x ...]
    const bool tl = p.prof != nullptr && blockIdx.y == 1 && blockIdx.x < 16 && threadIdx.x == 0;
    const uint32_t tl0 = clk32();
    unsigned long long* tlo = p.prof + 128 + 8 * blockIdx.x;
#define TL_MARK(k) do { if (tl) tlo[k] = clk32() - tl0; } while (0)
#else
#define TL_MARK(k) do { } while (0)
#endif

    if (warp == TMA_WARP && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == MMA_WARP && lane == 0) {
        mbar_init(q_full, 1);
        st_shared_u32(turn_word, 0u);
        for (int s = 0; s < SLOTS; ++s) {
            mbar_init(kv_full(s), 1);
            mbar_init(kv_empty(s), 1);
        }
        for (int i = 0; i < QT; ++i) {
            mbar_init(s_full(i), 1);
            for (int c = 0; c < 2; ++c) mbar_init(p_full(i, c), SOFTMAX_WARPS / QT);      // one arrival per softmax warp
            mbar_init(o_full(i), 1);
        }
        fence_barrier_init();
    }
    if (warp == ALLOC_WARP) {
        tmem_alloc(tmem_slot, Cfg<QT>::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    TL_MARK(0);

    if (warp >= SOFTMAX_WARPS) {
        if (warp == TMA_WARP && elect_one()) {
            // ------------------------------ TMA producer ------------------------------
            mbar_expect_tx(q_full, QT * TILE_BYTES);
#pragma unroll
            for (int i = 0; i < QT; ++i) {
                tma_load_2d(q_smem + i * TILE_BYTES, &tmQ, q_full, head * HD, q_row0 + i * BQ);
                tma_load_2d(q_smem + i * TILE_BYTES + HALF_BYTES, &tmQ, q_full, head * HD + 64, q_row0 + i * BQ);
            }
#pragma unroll 1
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
        } else if (warp == MMA_WARP && elect_one()) {
            // ------------------------------ MMA issuer ------------------------------
            // Every mbarrier probe costs ~100 cycles of latency on this single thread, so the schedule is a fixed
            // ping-pong (an event-driven poller over both tiles measured 2x slower); the K/V waits sit at the top of
            // the step, off the P -> PV -> QK chain of either tile.
            auto issue_qk = [&](int i, uint32_t k_addr) {
                const uint32_t qa = q_smem + i * TILE_BYTES;
                const uint32_t d = tmem_base + i * 128;
#pragma unroll
                for (int kk = 0; kk < HD / 16; ++kk) {
                    const uint32_t off = (kk >> 2) * HALF_BYTES + (kk & 3) * 32;
                    umma_ss(d, make_smem_desc_sw128(qa + off, 16, 1024), make_smem_desc_sw128(k_addr + off, 16, 1024),
                            IDESC_QK, kk != 0 ? 1u : 0u);
                }
            };
            // O_i += P_i[:, keys of hand-over c] V[keys of hand-over c, :].  Group g of half h of the tile holds keys
            // [64h + 16g, 64h + 16g + 16), i.e. the 16-key MMA step kk = 4h + g; hand-over 0 = groups [0, HO0_GROUPS).
            auto issue_pv = [&](int i, uint32_t v_addr, bool accumulate, int c) {
                const uint32_t d = tmem_base + O_COL + i * 128;
                const uint32_t pa = tmem_base + i * 128;
                const int g0 = c == 0 ? 0 : HO0_GROUPS, g1 = c == 0 ? HO0_GROUPS : NS / GC;
                bool first = !accumulate;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                    for (int g = 0; g < NS / GC; ++g) {
                        if (g < g0 || g >= g1) continue;
                        const int kk = 4 * hh + g;
                        // V tile: kv rows of 128 B (64 d-columns) per box; 16 kv rows = 2048 B; second d-half at +16 KB
                        umma_ts(d, pa + kk * 8, make_smem_desc_sw128(v_addr + kk * 2048, HALF_BYTES, 1024), IDESC_PV,
                                first ? 0u : 1u);
                        first = false;
                    }
                }
            };
            auto slot_of = [&](int t) { return t % SLOTS; };
            auto phase_of = [&](int t) { return static_cast<uint32_t>((t / SLOTS) & 1); };

            mbar_wait(q_full, 0, 0x200);
            mbar_wait(kv_full(slot_of(0)), phase_of(0), 0x210);
            tc_fence_after();
#pragma unroll
            for (int i = 0; i < QT; ++i) {
                issue_qk(i, kv_smem + slot_of(0) * TILE_BYTES);
                tc_commit(s_full(i));
            }
            tc_commit(kv_empty(slot_of(0)));
#pragma unroll 1
            for (int j = 0; j < n_kv; ++j) {
                const int tv = 2 * j + 1, tk = 2 * j + 2;
                const uint32_t pph = j & 1;
                const bool more = j + 1 < n_kv;
                mbar_wait(kv_full(slot_of(tv)), phase_of(tv), 0x220);
                if (more) mbar_wait(kv_full(slot_of(tk)), phase_of(tk), 0x240);
                const uint32_t v_addr = kv_smem + slot_of(tv) * TILE_BYTES;
                const uint32_t k_addr = kv_smem + slot_of(tk) * TILE_BYTES;
#pragma unroll
                for (int i = 0; i < QT; ++i) {
                    mbar_wait(p_full(i, 0), pph, 0x230 + i * 8);
                    tc_fence_after();
                    issue_pv(i, v_addr, j > 0, 0);
                    mbar_wait(p_full(i, 1), pph, 0x231 + i * 8);
                    tc_fence_after();
                    issue_pv(i, v_addr, true, 1);
                    if (i == QT - 1) tc_commit(kv_empty(slot_of(tv)));
                    if (more) {
                        issue_qk(i, k_addr);
                        tc_commit(s_full(i));
                        if (i == QT - 1) tc_commit(kv_empty(slot_of(tk)));
                    } else {
                        tc_commit(o_full(i));
                    }
                }
            }
        }
    } else {
        // ------------------------------ softmax warps ------------------------------
        const int i = warp >> 3;                    // Q tile of this warp (warps 0-7 / 8-15)
        const int h = (warp >> 2) & 1;              // which 64 score columns of the row this thread owns
        const int quarter = warp & 3;               // TMEM lane quarter accessible to this warp
        const int r = quarter * 32 + lane;          // row within the tile
        const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t s_tmem = tmem_base + i * 128 + lane_sel + h * NS;            // my score columns
        const uint32_t p_tmem = tmem_base + i * 128 + lane_sel + h * (NS / 2);      // my P columns (bf16 pairs)
        const uint32_t o_tmem = tmem_base + O_COL + i * 128 + lane_sel;
        const int row = q_row0 + i * BQ + r;
        const float sl2 = p.scale_log2;
        const uint64_t sl2_2 = f2_pack(sl2, sl2);
        const int tail_valid = p.sk - (n_kv - 1) * BKV - h * NS;    // my valid columns in the last KV tile (may be <= 0)
        const uint32_t pair_bar = 3 + i * 4 + quarter;              // named barrier of the two warps that share my rows
        auto xchg_mine = [&](int par) { return xchg + (((par * QT + i) * 2 + h) * BQ + r) * 4; };
        auto xchg_other = [&](int par) { return xchg + (((par * QT + i) * 2 + (1 - h)) * BQ + r) * 4; };
        // Reference point of the stored exponentials (raw-score units), identical in the two threads of a row.  It is
        // the true row maximum of the first KV tile and moves only when a later tile's row maximum exceeds it by more
        // than REF_MARGIN (softmax is invariant to the reference; fp32 and bf16 share the exponent range, so a stale
        // reference costs no precision until it would overflow).
        float m = -INFINITY;
        float l = 0.f;                              // sum over MY columns

        if (TURNS && i == 1) named_bar_arrive(1, 2 * (BQ * 2));      // tile 0 takes the first turn
#ifdef WVD_ATTN_PROF
        const bool prof = p.prof != nullptr && blockIdx.x == 1 && blockIdx.y == 0 && lane == 0;
        uint32_t pc_wait = 0, pc_ld = 0, pc_turn = 0, pc_max = 0, pc_b = 0, pc_redo = 0, pc_x1 = 0, pc_x2 = 0, pt = 0;
#endif
#pragma unroll 1
        for (int j = 0; j < n_kv; ++j) {
#ifdef WVD_ATTN_PROF
            if (prof) pt = clk32();
#endif
            mbar_wait(s_full(i), j & 1, 0x300 + i);
            tc_fence_after();
            if (j == 0) TL_MARK(1);
            PROF_LAP(pc_wait);
            uint32_t s[NS];
            tmem_ld_32x32b_x32(s_tmem + 0, s + 0);
            tmem_ld_32x32b_x32(s_tmem + 32, s + 32);
            tc_wait_ld();
            PROF_LAP(pc_ld);
            if (j == n_kv - 1 && tail_valid < NS) {
#pragma unroll
                for (int c = 0; c < NS; ++c)
                    if (c >= tail_valid) s[c] = 0xff800000u;   // -inf
            }
            // Row maximum of the tile -- only a guard against overflow of the stale reference.  Each thread reduces its
            // 64 columns; the two threads of a row exchange through shared memory (double-buffered by step parity).
            // All of this sits BEFORE the tile's turn on the MUFU pipe, i.e. normally in the shadow of the other
            // tile's exponentials.  The pair barrier also orders the P stores below after BOTH threads' S loads.
            const float mine = row_max<NS, 0, NS>(s, m);
#ifdef WVD_ATTN_PROF
            if (prof) { const uint32_t t = clk32() + (__float_as_uint(mine) & 0u); pc_x1 += t - pt; pt = t; }
#endif
            st_shared_u32(xchg_mine(j & 1), __float_as_uint(mine));
            named_bar_sync(pair_bar, 64);
            const float mx = fmaxf(mine, __uint_as_float(ld_shared_volatile_u32(xchg_other(j & 1))));
#ifdef WVD_ATTN_PROF
            if (prof) { const uint32_t t = clk32() + (__float_as_uint(mx) & 0u); pc_x2 += t - pt; pt = t; }
#endif
            // first tile: m = -inf and the comparison is true.  Warp-uniform, identical in the two warps of the pair.
            if (__any_sync(0xffffffffu, (mx - m) * sl2 > REF_MARGIN)) {
                // Move the reference to mx.  Every PV of the previous steps has completed (s_full(j) was committed
                // after them); the h = 0 thread of the row rescales its O accumulator, and the tensor core cannot
                // touch O again before all eight warps of the tile have handed P over.
                if (j > 0) {
                    const float alpha = fast_exp2((m - mx) * sl2);
                    l *= alpha;
                    if (h == 0) {
#pragma unroll 1
                        for (int c = 0; c < 8; ++c) {
                            uint32_t o[16];
                            tmem_ld_32x32b_x16(o_tmem + c * 16, o);
                            tc_wait_ld();
#pragma unroll
                            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                            tmem_st_32x32b_x16(o_tmem + c * 16, o);
                        }
                        tc_wait_st();
                    }
                }
                m = mx;
#ifdef WVD_ATTN_PROF
                if (prof) ++pc_redo;
#endif
            }
            PROF_LAP(pc_max);
            // My tile's turn on the MUFU pipe.  BAR.SYNC is deferred-blocking: the warp only stops at the next
            // instruction that depends on barrier-protected memory, so the exponentials are made to depend on a
            // (zero) word read after it.
            float turn_zero = 0.f;
            if (TURNS) {
                named_bar_sync(1 + i, 2 * (BQ * 2));
                turn_zero = __uint_as_float(ld_shared_volatile_u32(turn_word));
            }
#ifdef WVD_ATTN_PROF
            if (prof) { const uint32_t t = clk32() + __float_as_uint(turn_zero); pc_turn += t - pt; pt = t; }
#endif
            // Exponentials in groups of GC columns, each group stored to TMEM as soon as it is packed (few live
            // registers).
            const float neg_m = -m * sl2 + turn_zero;
            const uint64_t negm_2 = f2_pack(neg_m, neg_m);
            float lsum = 0.f;
#pragma unroll
            for (int g = 0; g < NS / GC; ++g) {
                uint32_t pk[GC / 2];
                switch (g) {   // compile-time after unrolling
                    case 0: lsum += exp_chunk<NS, 0 * GC, 1 * GC, EMU_OF_4>(s, pk, sl2_2, negm_2); break;
                    case 1: lsum += exp_chunk<NS, 1 * GC, 2 * GC, EMU_OF_4>(s, pk, sl2_2, negm_2); break;
                    case 2: lsum += exp_chunk<NS, 2 * GC, 3 * GC, EMU_OF_4>(s, pk, sl2_2, negm_2); break;
                    default: lsum += exp_chunk<NS, 3 * GC, 4 * GC, EMU_OF_4>(s, pk, sl2_2, negm_2); break;
                }
                if (TURNS && g == RELEASE_GROUP && (i == 0 || j + 1 < n_kv)) named_bar_arrive(2 - i, 2 * (BQ * 2));   // the other tile's turn
                store_p<GC / 2>(p_tmem + g * (GC / 2), pk);
                if (g == HO0_GROUPS - 1 || g == NS / GC - 1) {
                    // hand-over; the other warp of my scheduler keeps the MUFU pipe busy meanwhile
                    tc_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(p_full(i, g == NS / GC - 1 ? 1 : 0));
                }
            }
            l += lsum;
            PROF_LAP(pc_b);
        }
#ifdef WVD_ATTN_PROF
        if (prof) {
            unsigned long long* o = p.prof + warp * 8;
            o[0] = pc_wait; o[1] = pc_ld; o[2] = pc_max; o[3] = pc_b; o[4] = pc_redo; o[5] = n_kv; o[6] = pc_turn; o[7] = (static_cast<unsigned long long>(pc_x1) << 32) | pc_x2;
        }
#endif

        // ------------------------------ epilogue: O / l -> global ------------------------------
        // row sum = my columns + the other thread's
        TL_MARK(2);
        st_shared_u32(xchg_mine(0), __float_as_uint(l));
        named_bar_sync(pair_bar, 64);
        const float inv_l = 1.0f / (l + __uint_as_float(ld_shared_volatile_u32(xchg_other(0))));
        mbar_wait(o_full(i), 0, 0x310 + i);
        tc_fence_after();
        TL_MARK(3);
        // Every MMA that reads this tile's Q has completed (o_full): the CTA's Q tiles are dead and stage the output.
        const int wrow0 = q_row0 + i * BQ + quarter * 32;
        store_o_warp_coalesced(o_tmem + h * (HD / 2), inv_l, q_smem + warp * 4096, lane, [&](int rr) -> __nv_bfloat16* {
            const int grow = wrow0 + rr;
            if (grow >= p.sq) return nullptr;
            __nv_bfloat16* base;
            if (p.rows_per_peer > 0) {
                const int dest = grow / p.rows_per_peer;
                base = p.out_peer[dest < WVD_MAX_PEERS ? dest : 0] + static_cast<long long>(grow - dest * p.rows_per_peer) * p.ldo;
            } else {
                base = p.out + static_cast<long long>(grow) * p.ldo;
            }
            return base + head * HD + h * (HD / 2);
        });
    }

    TL_MARK(4);
    tc_fence_before();
    __syncthreads();
    TL_MARK(5);
    if (warp == ALLOC_WARP) tmem_dealloc(tmem_base, Cfg<QT>::TMEM_COLS);
#ifdef WVD_ATTN_PROF
    if (tl) {
        unsigned long long g;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
        tlo[6] = g;
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        tlo[7] = smid;
    }
#endif
}

#ifdef WVD_ATTN_PROF
unsigned long long* g_prof_buffer = nullptr;      // developer builds only (-DWVD_ATTN_PROF): in-kernel phase counters
#endif
}  // namespace attn

int attn_read_diag(unsigned long long* out) {
    if (cudaMemcpyFromSymbol(out, g_diag, sizeof(unsigned long long) * 8) != cudaSuccess) return -1;
    unsigned long long zero[8] = {0};
    cudaMemcpyToSymbol(g_diag, zero, sizeof(zero));
    return 0;
}

}  // namespace wvd

#ifdef WVD_ATTN_PROF
extern "C" __attribute__((visibility("default"))) int wvd_debug_attention_profile(unsigned long long* device_buf) {
    wvd::attn::g_prof_buffer = device_buf;
    return WVD_OK;
}
#endif

namespace wvd {
int attention_pair_launch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                          void* const* out_peers, int world, int64_t rows_per_peer, int64_t ldo, int num_heads,
                          int64_t sq, int64_t sk, float scale, cudaStream_t st);
int attention_cg2_launch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                         void* const* out_peers, int world, int64_t rows_per_peer, int64_t ldo, int num_heads,
                         int64_t sq, int64_t sk, float scale, cudaStream_t st);
int attention_cg2p_launch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                          void* const* out_peers, int world, int64_t rows_per_peer, int64_t ldo, int num_heads,
                          int64_t sq, int64_t sk, float scale, cudaStream_t st);
namespace attn {
static int launch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                  void* const* out_peers, int world, int64_t rows_per_peer, int64_t ldo, int num_heads, int64_t sq,
                  int64_t sk, int head_dim, float scale, int which, wvd_stream_t stream) {
    WVD_REQUIRE(q && k && v && (out || out_peers), "wvd_attention_fwd: null pointer");
    WVD_REQUIRE(which >= WVD_ATTN_AUTO && which <= WVD_ATTN_CG2_PERSISTENT, "wvd_attention_fwd: bad kernel selector %d", which);
    WVD_REQUIRE(head_dim == HD, "wvd_attention_fwd: head_dim must be 128 (got %d)", head_dim);
    WVD_REQUIRE(num_heads > 0 && num_heads <= 65535, "wvd_attention_fwd: bad num_heads %d", num_heads);
    WVD_REQUIRE(sq > 0 && sk > 0 && sq < (1ll << 31) && sk < (1ll << 31), "wvd_attention_fwd: bad sequence lengths sq=%lld sk=%lld", (long long)sq, (long long)sk);
    const int64_t width = (int64_t)num_heads * head_dim;
    WVD_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && ldq >= width && ldk >= width && ldv >= width && ldo >= width,
                "wvd_attention_fwd: leading dims must be multiples of 8 and >= heads*128");
    WVD_REQUIRE(((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0) && ((uintptr_t)out % 16 == 0),
                "wvd_attention_fwd: pointers must be 16-byte aligned");
    CUtensorMap tmQ, tmK, tmV;
    int rc = get_tensor_map_bf16(&tmQ, q, (uint64_t)sq, (uint64_t)width, (uint64_t)ldq, BQ);
    if (rc) return rc;
    rc = get_tensor_map_bf16(&tmK, k, (uint64_t)sk, (uint64_t)width, (uint64_t)ldk, BKV);
    if (rc) return rc;
    rc = get_tensor_map_bf16(&tmV, v, (uint64_t)sk, (uint64_t)width, (uint64_t)ldv, BKV);
    if (rc) return rc;
    Params p;
    p.out = (__nv_bfloat16*)out;
    p.ldo = ldo;
    p.rows_per_peer = 0;
    for (int r = 0; r < WVD_MAX_PEERS; ++r) p.out_peer[r] = nullptr;
    if (out_peers != nullptr) {
        WVD_REQUIRE(world >= 1 && world <= WVD_MAX_PEERS && rows_per_peer > 0 && rows_per_peer < (1ll << 31) &&
                    rows_per_peer * world >= sq, "wvd_attention_fwd_scatter: bad world / rows_per_peer");
        for (int r = 0; r < world; ++r) {
            WVD_REQUIRE(out_peers[r] && ((uintptr_t)out_peers[r] % 16 == 0), "wvd_attention_fwd_scatter: bad output pointer of rank %d", r);
            p.out_peer[r] = (__nv_bfloat16*)out_peers[r];
        }
        p.rows_per_peer = (int)rows_per_peer;
    }
    p.sq = (int)sq;
    p.sk = (int)sk;
    p.n_kv = (int)((sk + BKV - 1) / BKV);
    p.scale_log2 = scale * 1.4426950408889634f;
#ifdef WVD_ATTN_PROF
    p.prof = g_prof_buffer;
#else
    p.prof = nullptr;
#endif
    static unsigned long long configured = 0;
    if (first_use_on_current_device(&configured)) {
        WVD_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<2>::SMEM_BYTES));
        WVD_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<1>::SMEM_BYTES));
        WVD_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<0, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    cudaStream_t st = (cudaStream_t)stream;
    // Long key sequences (self-attention) go to the cta_group::2 kernel (attention_cg2_sm100.cu: one Q tile per CTA,
    // triple-buffered S, one M = 256 MMA stream per CTA pair with K/V split over the pair); its cta_group::1 predecessor
    // (attention_pair_sm100.cu, K/V multicast to both CTAs) stays selectable for A/B.  Short ones (the 512-token text cross-attention: 4 KV steps,
    // prologue-dominated) stay on the two-tile kernel of this file.  `which` is an explicit ARGUMENT (the parity
    // tests run both kernels on the same inputs): the library keeps no mutable selection state.
    if (which == WVD_ATTN_CG2_PERSISTENT || (which == WVD_ATTN_AUTO && sk >= 2048))
        return attention_cg2p_launch(q, ldq, k, ldk, v, ldv, out, out_peers ? (void* const*)p.out_peer : nullptr, world,
                                     rows_per_peer, ldo, num_heads, sq, sk, scale, st);
    if (which == WVD_ATTN_CG2)
        return attention_cg2_launch(q, ldq, k, ldk, v, ldv, out, out_peers ? (void* const*)p.out_peer : nullptr, world,
                                    rows_per_peer, ldo, num_heads, sq, sk, scale, st);
    if (which == WVD_ATTN_PAIR)
        return attention_pair_launch(q, ldq, k, ldk, v, ldv, out, out_peers ? (void* const*)p.out_peer : nullptr, world,
                                     rows_per_peer, ldo, num_heads, sq, sk, scale, st);
    // Short key sequences: the one-tile shape, two CTAs per SM (see Cfg); otherwise the 256-row CTA.
    if (which == WVD_ATTN_ONE_TILE || (which == WVD_ATTN_AUTO && sk <= 1024)) {
        dim3 grid((unsigned)((sq + BQ - 1) / BQ), (unsigned)num_heads);
        attention_fwd_kernel<0, 1><<<grid, Cfg<1>::NUM_THREADS, Cfg<1>::SMEM_BYTES, st>>>(tmQ, tmK, tmV, p);
    } else {
        dim3 grid((unsigned)((sq + 2 * BQ - 1) / (2 * BQ)), (unsigned)num_heads);
        attention_fwd_kernel<0, 2><<<grid, Cfg<2>::NUM_THREADS, Cfg<2>::SMEM_BYTES, st>>>(tmQ, tmK, tmV, p);
    }
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}
}  // namespace attn
}  // namespace wvd

extern "C" __attribute__((visibility("default"))) int wvd_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                 void* out, int64_t ldo, int num_heads, int64_t sq, int64_t sk, int head_dim,
                                 float scale, wvd_stream_t stream) {
    WVD_REQUIRE(out, "wvd_attention_fwd: null pointer");
    return wvd::attn::launch(q, ldq, k, ldk, v, ldv, out, nullptr, 1, 0, ldo, num_heads, sq, sk, head_dim, scale,
                             WVD_ATTN_AUTO, stream);
}

// Resident CTAs per SM of the short-key kernels (cudaOccupancyMaxActiveBlocksPerMultiprocessor): the one-tile shape is
// built so that two fit (registers, shared memory, TMEM columns); a build that loses that property loses its point.
extern "C" __attribute__((visibility("default"))) int wvd_debug_attention_resident_ctas(int which) {
    using namespace wvd::attn;
    int n = 0;
    if (which == WVD_ATTN_ONE_TILE) {
        cudaFuncSetAttribute(attention_fwd_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<1>::SMEM_BYTES);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, attention_fwd_kernel<0, 1>, Cfg<1>::NUM_THREADS, Cfg<1>::SMEM_BYTES) != cudaSuccess) return WVD_ERR_CUDA;
    } else {
        cudaFuncSetAttribute(attention_fwd_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<2>::SMEM_BYTES);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, attention_fwd_kernel<0, 2>, Cfg<2>::NUM_THREADS, Cfg<2>::SMEM_BYTES) != cudaSuccess) return WVD_ERR_CUDA;
    }
    return n;
}

// The same contraction on an explicitly named kernel (parity tests run both kernels on the same inputs).
extern "C" __attribute__((visibility("default"))) int wvd_attention_fwd_select(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                        void* out, int64_t ldo, int num_heads, int64_t sq, int64_t sk, int head_dim,
                                        float scale, int which, wvd_stream_t stream) {
    WVD_REQUIRE(out, "wvd_attention_fwd_select: null pointer");
    return wvd::attn::launch(q, ldq, k, ldk, v, ldv, out, nullptr, 1, 0, ldo, num_heads, sq, sk, head_dim, scale, which, stream);
}

// Ulysses return trip fused into the attention epilogue: out_ptrs[r] is rank r's (rows_per_peer, ldo) output buffer
// (peer pointers); query row t is stored at out_ptrs[t / rows_per_peer] + (t % rows_per_peer) * ldo + col_offset.
extern "C" __attribute__((visibility("default"))) int wvd_attention_fwd_scatter(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                         int64_t ldv, void* const* out_ptrs, int64_t ldo, int64_t rows_per_peer,
                                         int64_t col_offset, int world, int num_heads, int64_t sq, int64_t sk,
                                         int head_dim, float scale, int which, wvd_stream_t stream) {
    WVD_REQUIRE(out_ptrs && world >= 1 && world <= WVD_MAX_PEERS, "wvd_attention_fwd_scatter: bad peers");
    WVD_REQUIRE(col_offset >= 0 && col_offset % 8 == 0 && col_offset + (int64_t)num_heads * head_dim <= ldo,
                "wvd_attention_fwd_scatter: bad col_offset");
    void* shifted[WVD_MAX_PEERS];
    for (int r = 0; r < world; ++r) shifted[r] = out_ptrs[r] ? (void*)((__nv_bfloat16*)out_ptrs[r] + col_offset) : nullptr;
    // the leading-dimension check of launch() is against the local head count only
    return wvd::attn::launch(q, ldq, k, ldk, v, ldv, nullptr, shifted, world, rows_per_peer, ldo, num_heads, sq, sk, head_dim,
                             scale, which, stream);
}
