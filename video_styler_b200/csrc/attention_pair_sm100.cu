// K8b: self-attention forward for head_dim 128 on tcgen05 / TMEM (sm_100a), CTA-pair variant for long sequences.
//
//   out[s, h] = softmax(q_h k_h^T * scale) v_h          non-causal, no mask, no dropout
//   (replaces flash_attention(), diffsynth/models/wan_video_dit.py:28-61, for the ~30k-76k-token self-attention)
//
// Why a second kernel: in attention_sm100.cu (two Q tiles per CTA, S and P sharing TMEM columns) every tile runs the
// chain  softmax(j) -> PV(j) -> QK^T(j+1) -> softmax(j+1)  serially, and both the tensor pipe and the MUFU pipe idle
// ~40 % of the time (profiles/r1_attention_c3_v3.txt).  Here one CTA owns ONE 128-row Q tile, which leaves TMEM room to
// TRIPLE-BUFFER S: QK^T(j+2) is issued while softmax(j) is still running (into the buffer PV(j-1) has just released), so
// the softmax warps stream continuously and the tensor core is never on their critical path.  TMEM map (all 512
// columns): S_0 [0,128) | S_1 [128,256) | S_2 [256,384) | O [384,512); P(j) aliases S_(j%3)[0,64).
// One Q tile per SM would double the K/V traffic out of L2 (the binding resource at 148 SMs), so the two CTAs of a
// 2-CTA cluster (adjacent Q tiles of the same head) SHARE every K/V tile: each loads one half (64 keys) and TMA
// multicasts it into both CTAs' shared memory.  A slot is reused only when both CTAs have consumed it: each MMA issuer's
// tcgen05.commit is multicast to the slot's "free" barrier of BOTH CTAs.
// The single MMA-issuing thread turned out to be the last bound (every mbarrier probe costs it 100-400 cycles): K_{j+2}
// and V_j therefore travel in ONE ring slot behind ONE barrier, and barriers are indexed by the resource that limits how
// far a producer can run ahead (P hand-overs by S buffer) -- a barrier that can advance two phases before its consumer
// waits dead-locks on parity aliasing (tools/attn_stress.py, ~1 in 100 launches before the fix).
//   warps 0-3 / 4-7     softmax warpgroups: one query row per thread, warpgroup g owns the KV tiles j = g (mod 2)
//   warp 8 (1 thread)   TMA producer: Q tile once, then my 64-key half of (K_{j+2}, V_j) per step, multicast to the
//                       pair, ring of 3 pair slots (192 KB)
//   warp 9 (1 thread)   MMA issuer: S_(j+2)%3 = Q K_{j+2}^T (SS) issued AHEAD of O += P_j V_j (TS); owns the TMEM allocation
// Compile-time variants kept for A/B: WVD_ATTN2_PAIRSLOTS=0 (separate K / V slots and barriers),
// WVD_ATTN2_QTMEM (Q in TMEM, two S buffers), WVD_ATTN2_HO0 (keys in the first P hand-over / 16).  All of them compute
// the same result; the round-1 timing experiments that did not have been removed from the source.
#include <math.h>
#include <stdlib.h>

#include "host_utils.h"
#include "ptx.cuh"
#include "softmax_math.cuh"

namespace wvd {
#ifdef WVD_ATTN_PROF
namespace attn { extern unsigned long long* g_prof_buffer; }
#endif
namespace attn2 {

using attn::exp_chunk;
using attn::row_max;
using attn::store_p;

constexpr int BQ = 128, BKV = 128, HD = 128;
constexpr int GC = 16;                        // columns per exp2 / store group
#ifndef WVD_ATTN2_HO0
#define WVD_ATTN2_HO0 6
#endif
constexpr int HO0_GROUPS = WVD_ATTN2_HO0;     // groups of 16 keys in the first hand-over of P (8 = a single hand-over)
constexpr int TILE_BYTES = 128 * 128 * 2;     // 32 KB
constexpr int HALF_BYTES = TILE_BYTES / 2;    // one 64-column TMA box of 128 rows
constexpr int SLOTS = 6;                      // K/V ring: 6 x 32 KB (= 3 pair slots) + 32 KB of Q = 224 KB of shared memory
#ifdef WVD_ATTN2_QTMEM
// Variant: Q resident in TMEM (TS-mode QK^T: only K is read from shared memory, a third less operand traffic) at the
// price of the third S buffer: S_0 [0,128) | S_1 [128,256) | O [256,384) | Q [384,448) (bf16 pairs).
constexpr int SBUF = 2;
constexpr bool QTMEM = true;
#else
constexpr int SBUF = 3;                       // S buffers in TMEM
constexpr bool QTMEM = false;
#endif
#ifndef WVD_ATTN2_PAIRSLOTS
#define WVD_ATTN2_PAIRSLOTS 1
#endif
// K_{j+2} and V_j share one ring slot and ONE barrier: the MMA issuer's per-tile serial time (mbarrier probes cost
// ~100-200 cycles each on that single thread) is what bounds the kernel, so it waits once per step for its operands.
constexpr bool PAIRSLOTS = WVD_ATTN2_PAIRSLOTS != 0 && SBUF == 3;
constexpr int O_COL = SBUF * 128;             // first TMEM column of the O accumulator
constexpr int Q_COL = 384;                    // QTMEM only
constexpr int SOFTMAX_WARPS = 8, TMA_WARP = 8, MMA_WARP = 9;
constexpr int NUM_THREADS = 10 * 32;
constexpr int BAR_BYTES = 384;
constexpr int XCHG_BYTES = 3 * BQ * 4;        // m[row], l[warpgroup][row] fp32
constexpr int SMEM_BYTES = TILE_BYTES + SLOTS * TILE_BYTES + BAR_BYTES + XCHG_BYTES + 1024;
constexpr uint32_t IDESC_QK = make_idesc_bf16(128, 128, 0, 0);   // A = Q (K-major), B = K (K-major)
constexpr uint32_t IDESC_PV = make_idesc_bf16(128, 128, 0, 1);   // A = P (TMEM), B = V (MN-major)
constexpr float REF_MARGIN = 8.0f;

struct Params {
    const __nv_bfloat16* q;                   // QTMEM variant: the softmax threads read their Q rows from global memory
    long long ldq;
    __nv_bfloat16* out;
    long long ldo;
    __nv_bfloat16* out_peer[WVD_MAX_PEERS];   // Ulysses return trip fused into the epilogue (see attention_sm100.cu)
    int rows_per_peer;
    int sq, sk, n_kv;
    float scale_log2;
    unsigned long long* prof;   // optional device buffer (developer profiling, -DWVD_ATTN_PROF builds only)
};

#ifdef WVD_ATTN_PROF
#define PROF_LAP(acc) do { if (prof) { uint32_t t_; asm volatile("mov.u32 %0, %%clock;" : "=r"(t_)); (acc) += t_ - pt; pt = t_; } } while (0)
#else
#define PROF_LAP(acc) do { } while (0)
#endif

template <int EMU_OF_4>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
attention_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;     // same offset in both CTAs of the pair
    uint8_t* smem_gen = smem_raw + (smem_base - raw_u32);
    const uint32_t q_smem = smem_base;
    const uint32_t kv_smem = smem_base + TILE_BYTES;
    const uint32_t bar_base = kv_smem + SLOTS * TILE_BYTES;
    const uint32_t q_full = bar_base;
    auto kv_full = [&](int s) { return bar_base + 8 + s * 8; };                 // slot s holds a complete tile (both halves landed)
    auto kv_free = [&](int s) { return bar_base + 8 + (2 * SLOTS + s) * 8; };   // BOTH CTAs have finished reading slot s
    auto s_full = [&](int b) { return bar_base + 160 + b * 8; };                // S buffer b holds Q K^T
    // hand-over c of P of the tile in S buffer b is in TMEM.  Per BUFFER, not per warpgroup: with S triple-buffered the
    // softmax warps can hand over tiles j and j+2 before the MMA issuer has consumed tile j (two phases of a
    // per-warpgroup barrier -> parity aliasing -> deadlock); tile j+3 cannot be handed over before PV(j) was issued.
    auto p_full = [&](int b, int c) { return bar_base + 320 + (b * 2 + c) * 8; };
    const uint32_t o_full = bar_base + 216;
    auto pv_done = [&](int g) { return bar_base + 224 + g * 8; };              // PV of warpgroup g's latest tile (and every PV before it) has completed
    const uint32_t q_ready = bar_base + 368;                                     // QTMEM: Q is in TMEM
    const uint32_t tmem_slot = bar_base + 240;
    const uint32_t xchg = bar_base + BAR_BYTES;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (1 + SLOTS) * TILE_BYTES + 240);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int head = blockIdx.y;
    const int q_row0 = blockIdx.x * BQ;
    const int n_kv = p.n_kv;
    const uint32_t cta_rank = cluster_ctarank();

    if (warp == TMA_WARP && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == MMA_WARP && lane == 0) {
        mbar_init(q_full, 1);
        mbar_init(q_ready, SOFTMAX_WARPS / 2);
        for (int s = 0; s < SLOTS; ++s) {
            mbar_init(kv_full(s), 1);
            mbar_init(kv_free(s), 2);
        }
        for (int b = 0; b < SBUF; ++b) mbar_init(s_full(b), 1);
        for (int b = 0; b < SBUF; ++b)
            for (int c = 0; c < 2; ++c) mbar_init(p_full(b, c), SOFTMAX_WARPS / 2);     // one arrival per warp of the owning warpgroup
        mbar_init(o_full, 1);
        mbar_init(pv_done(0), 1);
        mbar_init(pv_done(1), 1);
        fence_barrier_init();
    }
    if (warp == MMA_WARP) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();        // the peer's barriers are initialised before any multicast / remote arrive reaches them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == TMA_WARP) {
        if (elect_one()) {
            // ------------------------------ TMA producer ------------------------------
            if (!QTMEM) {
                mbar_expect_tx(q_full, TILE_BYTES);
                tma_load_2d(q_smem, &tmQ, q_full, head * HD, q_row0);
                tma_load_2d(q_smem + HALF_BYTES, &tmQ, q_full, head * HD + 64, q_row0);
            }
            if (PAIRSLOTS) {
                // load q = j + 2 carries (K_q, V_{q-2}) into pair slot q % 3 (K at +0, V at +32 KB), one barrier for both
                for (int q = 0; q <= n_kv + 1; ++q) {
                    const bool has_k = q < n_kv, has_v = q >= 2;
                    if (!has_k && !has_v) continue;                          // q = 1 of a single-tile sequence
                    const int ps = q % 3;
                    if (q >= 3) mbar_wait(kv_free(ps), ((q / 3) - 1) & 1, 0x110 + ps);   // both CTAs are done with the slot
                    mbar_expect_tx(kv_full(ps), (has_k ? TILE_BYTES : 0) + (has_v ? TILE_BYTES : 0));
                    const uint32_t dst = kv_smem + ps * 2 * TILE_BYTES + cta_rank * (64 * 128);
                    if (has_k) {
                        const int row0 = q * BKV + static_cast<int>(cta_rank) * 64;
                        tma_load_2d_multicast(dst, &tmK, kv_full(ps), head * HD, row0, 0x3);
                        tma_load_2d_multicast(dst + HALF_BYTES, &tmK, kv_full(ps), head * HD + 64, row0, 0x3);
                    }
                    if (has_v) {
                        const int row0 = (q - 2) * BKV + static_cast<int>(cta_rank) * 64;
                        tma_load_2d_multicast(dst + TILE_BYTES, &tmV, kv_full(ps), head * HD, row0, 0x3);
                        tma_load_2d_multicast(dst + TILE_BYTES + HALF_BYTES, &tmV, kv_full(ps), head * HD + 64, row0, 0x3);
                    }
                }
            } else {
            // Tiles in the order the MMA issuer consumes them: K_0, K_1, then (K_{j+2}, V_j) for j = 0 .. n_kv-1.
            int t = 0;
            auto load = [&](const CUtensorMap* tm, int tile) {
                const int slot = t % SLOTS;
                if (t >= SLOTS)      // both CTAs' MMAs are done with the old tile (their commits are multicast to both CTAs)
                    mbar_wait(kv_free(slot), ((t / SLOTS) - 1) & 1, 0x110 + slot);
                mbar_expect_tx(kv_full(slot), TILE_BYTES);                  // 16 KB from me + 16 KB from the peer
                const uint32_t dst = kv_smem + slot * TILE_BYTES + cta_rank * (64 * 128);
                const int row0 = tile * BKV + static_cast<int>(cta_rank) * 64;
                tma_load_2d_multicast(dst, tm, kv_full(slot), head * HD, row0, 0x3);
                tma_load_2d_multicast(dst + HALF_BYTES, tm, kv_full(slot), head * HD + 64, row0, 0x3);
                ++t;
            };
            load(&tmK, 0);
            if (n_kv > 1) load(&tmK, 1);
#pragma unroll 1
            for (int j = 0; j < n_kv; ++j) {
                if (SBUF == 3) {                 // QK^T(j+2) is issued BEFORE PV(j)
                    if (j + 2 < n_kv) load(&tmK, j + 2);
                    load(&tmV, j);
                } else {                         // two S buffers: QK^T(j+2) reuses the buffer of P(j), after PV(j)
                    load(&tmV, j);
                    if (j + 2 < n_kv) load(&tmK, j + 2);
                }
            }
            }   // !PAIRSLOTS
        }
    } else if (warp == MMA_WARP) {
        if (elect_one()) {
            // ------------------------------ MMA issuer ------------------------------
            auto issue_qk = [&](int b, uint32_t k_addr) {
                const uint32_t d = tmem_base + b * 128;
#pragma unroll
                for (int kk = 0; kk < HD / 16; ++kk) {
                    const uint32_t off = (kk >> 2) * HALF_BYTES + (kk & 3) * 32;
                    if (QTMEM)
                        umma_ts(d, tmem_base + Q_COL + kk * 8, make_smem_desc_sw128(k_addr + off, 16, 1024), IDESC_QK,
                                kk != 0 ? 1u : 0u);
                    else
                        umma_ss(d, make_smem_desc_sw128(q_smem + off, 16, 1024), make_smem_desc_sw128(k_addr + off, 16, 1024),
                                IDESC_QK, kk != 0 ? 1u : 0u);
                }
            };
            // O += P[:, keys of hand-over c] V[keys of hand-over c, :]; hand-over 0 = the first HO0_GROUPS groups of 16 keys
            auto issue_pv = [&](int b, uint32_t v_addr, bool accumulate, int c) {
                const uint32_t d = tmem_base + O_COL;
                const uint32_t pa = tmem_base + b * 128;
                const int k0 = c == 0 ? 0 : HO0_GROUPS, k1 = c == 0 ? HO0_GROUPS : BKV / GC;
#pragma unroll
                for (int kk = 0; kk < BKV / GC; ++kk) {
                    if (kk < k0 || kk >= k1) continue;
                    umma_ts(d, pa + kk * 8, make_smem_desc_sw128(v_addr + kk * 2048, HALF_BYTES, 1024), IDESC_PV,
                            (accumulate || kk != 0) ? 1u : 0u);
                }
            };
            int t = 0;
            int slot = 0;
            auto next_tile = [&]() {
                slot = t % SLOTS;
                mbar_wait(kv_full(slot), (t / SLOTS) & 1, 0x200 + slot);
                ++t;
                return kv_smem + slot * TILE_BYTES;
            };
            if (QTMEM) mbar_wait(q_ready, 0, 0x211); else mbar_wait(q_full, 0, 0x210);
            tc_fence_after();
            if (PAIRSLOTS) {
                for (int q = 0; q < 2 && q < n_kv; ++q) {                    // K_0, K_1
                    mbar_wait(kv_full(q), 0, 0x200 + q);
                    tc_fence_after();
                    issue_qk(q, kv_smem + q * 2 * TILE_BYTES);
                    tc_commit(s_full(q));
                    tc_commit_multicast(kv_free(q), 0x3);
                }
                int b = 0;                                   // j % 3
                uint32_t b_round = 0;                        // j / 3
#pragma unroll 1
                for (int j = 0; j < n_kv; ++j) {
                    const int q = j + 2, ps = b == 0 ? 2 : b - 1;             // ps = q % 3 = (j + 2) % 3
                    mbar_wait(kv_full(ps), (q / 3) & 1, 0x200 + ps);         // K_{j+2} and V_j, one probe
                    tc_fence_after();
                    const uint32_t base = kv_smem + ps * 2 * TILE_BYTES;
                    if (j + 2 < n_kv) {
                        // QK^T two steps ahead, into the buffer whose P was consumed by PV(j-1) (issued in the last iteration)
                        issue_qk(ps, base);
                        tc_commit(s_full(ps));
                    }
                    mbar_wait(p_full(b, 0), b_round & 1, 0x220);
                    tc_fence_after();
                    issue_pv(b, base + TILE_BYTES, j > 0, 0);
                    if (HO0_GROUPS < BKV / GC) {
                        mbar_wait(p_full(b, 1), b_round & 1, 0x221);
                        tc_fence_after();
                        issue_pv(b, base + TILE_BYTES, true, 1);
                    }
                    tc_commit_multicast(kv_free(ps), 0x3);
                    tc_commit(pv_done(j & 1));
                    if (j + 1 == n_kv) tc_commit(o_full);
                    if (++b == SBUF) { b = 0; ++b_round; }
                }
            } else {
            for (int j0 = 0; j0 < 2 && j0 < n_kv; ++j0) {
                const uint32_t k_addr = next_tile();
                tc_fence_after();
                issue_qk(j0, k_addr);
                tc_commit(s_full(j0));
                tc_commit_multicast(kv_free(slot), 0x3);
            }
            int b = 0;                                   // j % 3
            uint32_t b_round = 0;                        // j / 3
#ifdef WVD_ATTN_PROF
            const bool prof = p.prof != nullptr && blockIdx.x == 2 && blockIdx.y == 0;
            uint32_t pc_k = 0, pc_v = 0, pc_p0 = 0, pc_p1 = 0, pc_issue = 0, pt = 0;
            if (prof) asm volatile("mov.u32 %0, %%clock;" : "=r"(pt));
#endif
#pragma unroll 1
            for (int j = 0; j < n_kv; ++j) {
                if (SBUF == 3 && j + 2 < n_kv) {
                    // QK^T two steps ahead, into the buffer whose P was consumed by PV(j-1) (issued in the last iteration)
                    const int b2 = b == 0 ? 2 : b - 1;   // (j + 2) % 3
                    const uint32_t k_addr = next_tile();
                    PROF_LAP(pc_k);
                    tc_fence_after();
                    issue_qk(b2, k_addr);
                    tc_commit(s_full(b2));
                    tc_commit_multicast(kv_free(slot), 0x3);
                    PROF_LAP(pc_issue);
                }
                const uint32_t v_addr = next_tile();
                const int v_slot = slot;
                PROF_LAP(pc_v);
                mbar_wait(p_full(b, 0), b_round & 1, 0x220);
                PROF_LAP(pc_p0);
                tc_fence_after();
                issue_pv(b, v_addr, j > 0, 0);
                PROF_LAP(pc_issue);
                if (HO0_GROUPS < BKV / GC) {
                    mbar_wait(p_full(b, 1), b_round & 1, 0x221);
                    PROF_LAP(pc_p1);
                    tc_fence_after();
                    issue_pv(b, v_addr, true, 1);
                    PROF_LAP(pc_issue);
                }
                tc_commit_multicast(kv_free(v_slot), 0x3);
                tc_commit(pv_done(j & 1));
                if (j + 1 == n_kv) tc_commit(o_full);
                if (SBUF == 2 && j + 2 < n_kv) {
                    // two S buffers: QK^T(j+2) goes into the buffer P(j) has just been read from (in-order tensor pipe)
                    const uint32_t k_addr = next_tile();
                    tc_fence_after();
                    issue_qk(b, k_addr);
                    tc_commit(s_full(b));
                    tc_commit_multicast(kv_free(slot), 0x3);
                }
                if (++b == SBUF) { b = 0; ++b_round; }
            }
#ifdef WVD_ATTN_PROF
            if (prof) {
                unsigned long long* o = p.prof + MMA_WARP * 8;
                o[0] = pc_k; o[1] = pc_v; o[2] = pc_p0; o[3] = pc_p1; o[4] = pc_issue; o[5] = n_kv;
            }
#endif
            }   // !PAIRSLOTS
        }
    } else {
        // ------------------------------ softmax warps ------------------------------
        // Warpgroup g (warps 4g .. 4g+3) owns the KV tiles j = g (mod 2); one query row per thread (all
        // 128 score columns).  The two threads of a row (same scheduler, warps w and w+4) work on consecutive KV
        // tiles half a period apart, so together they keep the MUFU pipe fed although a lone warp cannot (in-order
        // issue exposes the MUFU latency), and neither ever waits for the tensor core: S(j+2) is produced during tile j.  The only thing they share is the reference point m of the stored
        // exponentials, handed from the thread of tile j-1 to the thread of tile j through shared memory: it moves
        // only when the exact row maximum of a tile exceeds it by 2^REF_MARGIN (lazy rescale of O and l).
        const int g = warp >> 2;
        const int quarter = warp & 3;               // TMEM lane quarter accessible to this warp
        const int r = quarter * 32 + lane;
        const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t o_tmem = tmem_base + O_COL + lane_sel;
        const int row = q_row0 + r;
        const float sl2 = p.scale_log2;
        const uint64_t sl2_2 = f2_pack(sl2, sl2);
        const int tail_valid = p.sk - (n_kv - 1) * BKV;              // valid keys in the last KV tile (1..128)
        const uint32_t bar_mine = 1 + g * 4 + quarter;               // I arrive here once I have published m for my tile
        const uint32_t bar_other = 1 + (1 - g) * 4 + quarter;        // ... and wait here for the m of the tile before
        const uint32_t bar_pair = 9 + quarter;                       // both threads of the row (epilogue)
        const uint32_t m_addr = xchg + r * 4;
        const uint32_t l_addr = xchg + (BQ + g * BQ + r) * 4, l_other_addr = xchg + (BQ + (1 - g) * BQ + r) * 4;
        if (QTMEM && g == 0) {
            // my Q row -> TMEM (A operand of the TS-mode QK^T: lane = row, 32-bit column c = elements 2c, 2c+1)
            uint32_t qw[64];
            const uint4* src = reinterpret_cast<const uint4*>(p.q + static_cast<long long>(row) * p.ldq + head * HD);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const uint4 v = row < p.sq ? __ldg(src + c) : make_uint4(0u, 0u, 0u, 0u);
                qw[4 * c] = v.x; qw[4 * c + 1] = v.y; qw[4 * c + 2] = v.z; qw[4 * c + 3] = v.w;
            }
            tmem_st_32x32b_x32(tmem_base + Q_COL + lane_sel, qw);
            tmem_st_32x32b_x32(tmem_base + Q_COL + 32 + lane_sel, qw + 32);
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_ready);
        }
        float m_last = -INFINITY;     // the reference my l is expressed in
        float l = 0.f;                // sum over MY tiles

        int b = g;                    // j % 3 of my current tile
        uint32_t b_round = 0;         // j / 3
#ifdef WVD_ATTN_PROF
        const bool prof = p.prof != nullptr && blockIdx.x == 2 && blockIdx.y == 0 && lane == 0;
        uint32_t pc_wait = 0, pc_ld = 0, pc_max = 0, pc_sync = 0, pc_dec = 0, pc_exp = 0, pc_steps = 0, pt = 0;
#endif
#pragma unroll 1
        for (int j = g; j < n_kv; j += 2) {
            const uint32_t s_tmem = tmem_base + b * 128 + lane_sel;        // S buffer of tile j; P aliases its columns [0,64)
#ifdef WVD_ATTN_PROF
            if (prof) { asm volatile("mov.u32 %0, %%clock;" : "=r"(pt)); ++pc_steps; }
#endif
            mbar_wait(s_full(b), b_round & 1, 0x300 + b);
            tc_fence_after();
            PROF_LAP(pc_wait);
            uint32_t s[BKV];
            tmem_ld_32x32b_x32(s_tmem + 0, s + 0);
            tmem_ld_32x32b_x32(s_tmem + 32, s + 32);
            tmem_ld_32x32b_x32(s_tmem + 64, s + 64);
            tmem_ld_32x32b_x32(s_tmem + 96, s + 96);
            tc_wait_ld();
            PROF_LAP(pc_ld);
            if (j == n_kv - 1 && tail_valid < BKV) {
#pragma unroll
                for (int c = 0; c < BKV; ++c)
                    if (c >= tail_valid) s[c] = 0xff800000u;   // -inf
            }
            const float mx = row_max<BKV, 0, BKV>(s, -INFINITY);     // exact row maximum of this tile
#ifdef WVD_ATTN_PROF
            if (prof) { uint32_t t_; asm volatile("mov.u32 %0, %%clock;" : "=r"(t_)); t_ += __float_as_uint(mx) & 0u; pc_max += t_ - pt; pt = t_; }
#endif
            float m_prev = -INFINITY;
            if (j > 0) {
                // The thread of tile j-1 (other warpgroup, same scheduler) has published its m: producer / consumer
                // named barrier (bar.arrive by the publisher, bar.sync here).  It cannot be signalled twice before I
                // consume it: the publisher's next tile needs MY decision first.
                named_bar_sync(bar_other, 64);
                m_prev = __uint_as_float(ld_shared_volatile_u32(m_addr));
            }
#ifdef WVD_ATTN_PROF
            if (prof) { uint32_t t_; asm volatile("mov.u32 %0, %%clock;" : "=r"(t_)); t_ += __float_as_uint(m_prev) & 0u; pc_sync += t_ - pt; pt = t_; }
#endif
            const float m_new = ((mx - m_prev) * sl2 > REF_MARGIN) ? mx : m_prev;     // tile 0: m_prev = -inf -> mx
            st_shared_u32(m_addr, __float_as_uint(m_new));
            if (j + 1 < n_kv) named_bar_arrive(bar_mine, 64);
            // Rare path (warp-uniform): some row of this warp moves its reference, or my l is in an older reference
            // (m_last <= m_prev <= m_new, so one comparison covers both).
            if (__any_sync(0xffffffffu, m_new != m_last)) {
                if (j > 0 && __any_sync(0xffffffffu, m_new != m_prev)) {
                    // O holds tiles < j relative to m_prev and PV(j-1) may still be accumulating: wait for it, rescale
                    // my row.  Nobody else touches O meanwhile: PV(j) needs my P, and the thread of tile j+1 can only
                    // rescale after PV(j).  PV(j-1) belongs to the OTHER warpgroup's barrier, which can only be one
                    // phase away from what I expect: its previous tile j-3 completed before S(j) did, its next tile
                    // j+1 needs PV(j).
                    mbar_wait(pv_done(1 - g), ((j - 1) >> 1) & 1, 0x320);
                    tc_fence_after();
                    const float alpha = fast_exp2((m_prev - m_new) * sl2);       // 1 for the rows that did not move
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
                l *= fast_exp2((m_last - m_new) * sl2);       // 1 if unchanged; first own tile: l = 0, exp2(-inf) = 0
            }
            m_last = m_new;
            PROF_LAP(pc_dec);
            const float neg_m = -m_new * sl2;
            const uint64_t negm_2 = f2_pack(neg_m, neg_m);
            float lsum = 0.f;
#pragma unroll
            for (int q8 = 0; q8 < BKV / GC; ++q8) {
                uint32_t pk[GC / 2];
                switch (q8) {   // compile-time after unrolling
                    case 0: lsum += exp_chunk<BKV, 0 * GC, 1 * GC, EMU_OF_4>(s, pk, sl2_2, negm_2); break;
                    case 1: lsum += exp_chunk<BKV, 1 * GC, 2 * GC, EMU_OF_4>(s, pk, sl2_2, negm_2); break;
                    case 2: lsum += exp_chunk<BKV, 2 * GC, 3 * GC, EMU_OF_4>(s, pk, sl2_2, negm_2); break;
                    case 3: lsum += exp_chunk<BKV, 3 * GC, 4 * GC, EMU_OF_4>(s, pk, sl2_2, negm_2); break;
                    case 4: lsum += exp_chunk<BKV, 4 * GC, 5 * GC, EMU_OF_4>(s, pk, sl2_2, negm_2); break;
                    case 5: lsum += exp_chunk<BKV, 5 * GC, 6 * GC, EMU_OF_4>(s, pk, sl2_2, negm_2); break;
                    case 6: lsum += exp_chunk<BKV, 6 * GC, 7 * GC, EMU_OF_4>(s, pk, sl2_2, negm_2); break;
                    default: lsum += exp_chunk<BKV, 7 * GC, 8 * GC, EMU_OF_4>(s, pk, sl2_2, negm_2); break;
                }
                store_p<GC / 2>(s_tmem + q8 * (GC / 2), pk);
                if (q8 == HO0_GROUPS - 1 || q8 == BKV / GC - 1) {
                    tc_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(p_full(b, q8 == HO0_GROUPS - 1 ? 0 : 1));
                }
            }
            l += lsum;
            PROF_LAP(pc_exp);
            b += 2;
            if (b >= SBUF) { b -= SBUF; ++b_round; }
        }
#ifdef WVD_ATTN_PROF
        if (prof) {
            unsigned long long* o = p.prof + warp * 8;
            o[0] = pc_wait; o[1] = pc_ld; o[2] = pc_max; o[3] = pc_sync; o[4] = pc_dec; o[5] = pc_exp; o[6] = pc_steps;
        }
#endif

        // ------------------------------ epilogue: O / l -> global ------------------------------
        named_bar_sync(bar_pair, 64);                                // both threads of the row are past their last tile
        const float m_fin = __uint_as_float(ld_shared_volatile_u32(m_addr));
        l *= fast_exp2((m_last - m_fin) * sl2);                      // no own tile (n_kv = 1, g = 1): 0 * exp2(-inf) = 0
        st_shared_u32(l_addr, __float_as_uint(l));
        named_bar_sync(bar_pair, 64);
        const float inv_l = 1.0f / (l + __uint_as_float(ld_shared_volatile_u32(l_other_addr)));
        mbar_wait(o_full, 0, 0x310);
        tc_fence_after();
        __nv_bfloat16* orow;
        if (p.rows_per_peer > 0) {
            const int dest = row / p.rows_per_peer;
            orow = p.out_peer[dest < WVD_MAX_PEERS ? dest : 0] + static_cast<long long>(row - dest * p.rows_per_peer) * p.ldo;
        } else {
            orow = p.out + static_cast<long long>(row) * p.ldo;
        }
        orow += head * HD + g * (HD / 2);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(o_tmem + g * (HD / 2) + c * 32, o);
            tc_wait_ld();
            if (row < p.sq) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(o[q4 * 8 + 0]) * inv_l, __uint_as_float(o[q4 * 8 + 1]) * inv_l);
                    u.y = pack_bf16x2(__uint_as_float(o[q4 * 8 + 2]) * inv_l, __uint_as_float(o[q4 * 8 + 3]) * inv_l);
                    u.z = pack_bf16x2(__uint_as_float(o[q4 * 8 + 4]) * inv_l, __uint_as_float(o[q4 * 8 + 5]) * inv_l);
                    u.w = pack_bf16x2(__uint_as_float(o[q4 * 8 + 6]) * inv_l, __uint_as_float(o[q4 * 8 + 7]) * inv_l);
                    *reinterpret_cast<uint4*>(orow + c * 32 + q4 * 8) = u;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();        // neither CTA leaves while the other may still multicast into it or arrive on its barriers
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, 512);
}

}  // namespace attn2

int attn_pair_read_diag(unsigned long long* out) {
    if (cudaMemcpyFromSymbol(out, g_diag, sizeof(unsigned long long) * 8) != cudaSuccess) return -1;
    unsigned long long zero[8] = {0};
    cudaMemcpyToSymbol(g_diag, zero, sizeof(zero));
    return 0;
}

// Launch the CTA-pair kernel.  Same contract as wvd::attn::launch (attention_sm100.cu), which validates the arguments.
int attention_pair_launch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                          void* const* out_peers, int world, int64_t rows_per_peer, int64_t ldo, int num_heads,
                          int64_t sq, int64_t sk, float scale, cudaStream_t st) {
    using namespace attn2;
    const int64_t width = (int64_t)num_heads * HD;
    CUtensorMap tmQ, tmK, tmV;
    int rc = get_tensor_map_bf16(&tmQ, q, (uint64_t)sq, (uint64_t)width, (uint64_t)ldq, BQ);
    if (rc) return rc;
    rc = get_tensor_map_bf16(&tmK, k, (uint64_t)sk, (uint64_t)width, (uint64_t)ldk, BKV / 2);
    if (rc) return rc;
    rc = get_tensor_map_bf16(&tmV, v, (uint64_t)sk, (uint64_t)width, (uint64_t)ldv, BKV / 2);
    if (rc) return rc;
    Params p;
    p.q = (const __nv_bfloat16*)q;
    p.ldq = ldq;
    p.out = (__nv_bfloat16*)out;
    p.ldo = ldo;
    p.rows_per_peer = 0;
    for (int r = 0; r < WVD_MAX_PEERS; ++r) p.out_peer[r] = nullptr;
    if (out_peers != nullptr) {
        for (int r = 0; r < world; ++r) p.out_peer[r] = (__nv_bfloat16*)out_peers[r];
        p.rows_per_peer = (int)rows_per_peer;
    }
    p.sq = (int)sq;
    p.sk = (int)sk;
    p.n_kv = (int)((sk + BKV - 1) / BKV);
    p.scale_log2 = scale * 1.4426950408889634f;
#ifdef WVD_ATTN_PROF
    p.prof = attn::g_prof_buffer;
#else
    p.prof = nullptr;
#endif
    static unsigned long long configured = 0;
    if (first_use_on_current_device(&configured))
        WVD_CHECK_CUDA(cudaFuncSetAttribute(attention_pair_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    unsigned q_tiles = (unsigned)((sq + BQ - 1) / BQ);
    q_tiles = (q_tiles + 1u) & ~1u;          // whole CTA pairs; a surplus CTA computes rows >= sq and stores nothing
    dim3 grid(q_tiles, (unsigned)num_heads);
    attention_pair_kernel<0><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmQ, tmK, tmV, p);
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}

}  // namespace wvd
