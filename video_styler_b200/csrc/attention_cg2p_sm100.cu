// K8d: self-attention forward for head_dim 128 on tcgen05 / TMEM (sm_100a): PERSISTENT CTA-pair kernel (cta_group::2).
//
//   out[s, h] = softmax(q_h k_h^T * scale) v_h          non-causal, no mask, no dropout
//   (replaces flash_attention(), diffsynth/models/wan_video_dit.py:28-61, for the ~30k-76k-token self-attention)
//
// The inner pipeline is the one of attention_cg2_sm100.cu (one 128-row Q tile per CTA, one tcgen05.mma cta_group::2 stream
// with M = 256 per CTA pair, K split over the pair by keys and V by head-dim columns, S triple-buffered in TMEM with P(j)
// over S(j), QK^T two steps ahead of PV, two softmax warpgroups alternating KV steps).  What changes is the OUTSIDE of it.
// The non-persistent kernel launches one CTA pair per (256 query rows, head) work item; its in-kernel timeline
// (profiles/r2_attention_timelines.txt) shows ~5,400 cycles from CTA entry to the first S (every CTA of a wave pulls its Q
// tile from HBM at the same moment: ~16 B/cycle/SM) and ~4,400 cycles of epilogue (every CTA stores at the same moment),
// i.e. ~10,000 of the ~178,000 cycles of a c3 work item with the tensor core idle.  Here one CTA pair per SM pair stays
// resident and walks a list of work items; the pipeline never drains:
//   * K/V loads, S buffers, P hand-overs and PV/QK^T issue run over ONE global sequence of KV steps G = item * n_kv + j;
//     ring-slot load L carries K of step L and V of step L-2 across item boundaries (the heads may differ);
//   * Q is triple-buffered: Q of item k+1 is requested at the start of item k (its buffer was last used, as Q and as
//     output staging, by item k-2), so even the 4-step items of the 512-key cross-attention get their Q a whole item ahead;
//   * QK^T of the first two steps of item k+1 is issued during the last two steps of item k, so S(k+1, 0) is waiting when
//     the softmax warps come back from the epilogue of item k;
//   * the epilogue of item k (TMEM -> bf16 -> the dead Q buffer of item k as staging -> 128-byte row segments to global)
//     runs on the softmax warps while the tensor core already works on item k+1; the first PV of item k+1 (which
//     overwrites O) waits for one "O drained" barrier per item.
// Work items are dealt round-robin (item = cluster + k * clusters, head-major so that concurrently running items share
// K/V in L2); all items cost the same.  For very short key sequences (n_kv < 4: two steps ahead would cross more than one
// item boundary) the same kernel runs with one item per CTA pair, i.e. as the non-persistent kernel.
// Only the LEADER CTA (cluster rank 0) issues MMAs.  Both CTAs' TMA loads complete on the leader's barriers
// (cp.async.bulk.tensor .cta_group::2), both CTAs' softmax warps hand P over on the leader's barriers (remote arrive),
// tcgen05.commit is multicast to the S-full / slot-free / PV-done / O-full barriers of both CTAs.
//   warps 0-3 / 4-7     softmax warpgroups: one query row per thread, warpgroup g owns the global KV steps G = g (mod 2)
//   warp 8 (1 thread)   TMA producer: my Q tiles, my half of (K_L, V_{L-2}) per load, one ring slot + ONE barrier
//   warp 9 (1 thread)   MMA issuer (leader CTA only); the warp owns the pair-wide TMEM allocation in both CTAs
#include <math.h>

#include "host_utils.h"
#include "ptx.cuh"
#include "softmax_math.cuh"

namespace wvd {
namespace attn4 {

using attn::exp_chunk;
using attn::row_max;
using attn::store_p;

constexpr int BQ = 128, BKV = 128, HD = 128;
constexpr int GC = 16;                        // columns per exp2 / store group
constexpr int HO0_GROUPS = 6;                 // groups of 16 keys in the first hand-over of P
constexpr int Q_BYTES = 128 * 128 * 2;        // 32 KB: two 64-column boxes of 128 rows
constexpr int QBOX_BYTES = Q_BYTES / 2;
constexpr int KHALF_BYTES = 64 * 128 * 2;     // my 64 keys x 128 d: two 64-column boxes of 64 rows (8 KB each)
constexpr int KBOX_BYTES = KHALF_BYTES / 2;
constexpr int VHALF_BYTES = 128 * 64 * 2;     // 128 keys x my 64 d columns: one box
constexpr int SLOT_BYTES = KHALF_BYTES + VHALF_BYTES;      // 32 KB: (K_L, V_{L-2}) halves
#ifndef WVD_CG2P_SLOTS
#define WVD_CG2P_SLOTS 4
#endif
constexpr int SLOTS = WVD_CG2P_SLOTS;         // 5 also fits (232,320 of 232,448 bytes)
static_assert(SLOTS <= 5, "barrier layout holds at most 5 ring slots");
// shared memory: 3 Q buffers + 4 slots = 224 KB + barriers / exchange = 232,320 of the 232,448 bytes a CTA may have
constexpr int QBUFS = 3;                      // Q(k+1) is requested at the START of item k: its buffer was last used by item k-2
constexpr int SBUF = 3;                       // S buffers in TMEM
constexpr int O_COL = SBUF * 128;             // first TMEM column of the O accumulator
constexpr int SOFTMAX_WARPS = 8, TMA_WARP = 8, MMA_WARP = 9;
constexpr int NUM_THREADS = 10 * 32;
constexpr int BAR_BYTES = 384;
constexpr int XCHG_BYTES = 3 * BQ * 4;        // m[row], l[warpgroup][row] fp32
constexpr int SMEM_BYTES = QBUFS * Q_BYTES + SLOTS * SLOT_BYTES + BAR_BYTES + XCHG_BYTES + 1024;
constexpr int MIN_KV_PERSISTENT = 4;         // two steps ahead must cross at most ONE item boundary, and the Q prefetch needs >= 2 steps
constexpr uint32_t IDESC_QK = make_idesc_bf16(256, 128, 0, 0);   // A = Q (K-major), B = K (K-major), M = 256 over the pair
constexpr uint32_t IDESC_PV = make_idesc_bf16(256, 128, 0, 1);   // A = P (TMEM), B = V (MN-major)
constexpr float REF_MARGIN = 8.0f;

struct Params {
    __nv_bfloat16* out;
    long long ldo;
    __nv_bfloat16* out_peer[WVD_MAX_PEERS];   // Ulysses return trip fused into the epilogue (see attention_sm100.cu)
    int rows_per_peer;
    int sq, sk, n_kv;
    int n_qpairs, n_items, n_clusters;        // work items = (pair of Q tiles, head), head-major
    float scale_log2;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
attention_cg2p_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;     // same offset in both CTAs of the pair
    uint8_t* smem_gen = smem_raw + (smem_base - raw_u32);
    const uint32_t q_smem = smem_base;
    const uint32_t kv_smem = smem_base + QBUFS * Q_BYTES;
    const uint32_t bar_base = kv_smem + SLOTS * SLOT_BYTES;
    auto q_full = [&](int qb) { return bar_base + qb * 8; };                     // LEADER: both CTAs' Q tiles have landed in buffer qb
    auto q_free = [&](int qb) { return bar_base + 24 + qb * 8; };                // MINE: my 8 epilogue warps are done with buffer qb (Q reads + staging)
    auto kv_full = [&](int s) { return bar_base + 48 + s * 8; };                 // LEADER: both halves of slot s have landed
    auto kv_free = [&](int s) { return bar_base + 88 + s * 8; };                 // the MMAs reading slot s have completed
    auto s_full = [&](int b) { return bar_base + 128 + b * 8; };                 // S buffer b holds Q K^T
    // LEADER: hand-over c of P of the step in S buffer b is in TMEM of BOTH CTAs.  Per BUFFER, not per warpgroup (parity
    // aliasing, see attention_cg2_sm100.cu).
    auto p_full = [&](int b, int c) { return bar_base + 160 + (b * 2 + c) * 8; };
    auto pv_done = [&](int g) { return bar_base + 208 + g * 8; };               // PV of the latest step of parity g (and every PV before it) has completed
    const uint32_t o_full = bar_base + 224;                                      // every PV of the current item has completed
    const uint32_t o_free = bar_base + 232;                                      // LEADER: all 16 softmax warps of the pair have read O out of TMEM
    const uint32_t tmem_slot = bar_base + 240;
    const uint32_t xchg = bar_base + BAR_BYTES;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + QBUFS * Q_BYTES + SLOTS * SLOT_BYTES + 240);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_kv = p.n_kv;
    const uint32_t rank = cluster_ctarank();
    const int cluster = blockIdx.x >> 1;
    // my work items: cluster, cluster + n_clusters, ...
    const int n_my = (p.n_items - cluster + p.n_clusters - 1) / p.n_clusters;
    const int total = n_my * n_kv;                                // global KV steps of this CTA pair
    auto item_head = [&](int k) { return (cluster + k * p.n_clusters) / p.n_qpairs; };
    auto item_row0 = [&](int k) { return (((cluster + k * p.n_clusters) % p.n_qpairs) * 2 + static_cast<int>(rank)) * BQ; };

    if (warp == TMA_WARP && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == MMA_WARP && lane == 0) {
        for (int qb = 0; qb < QBUFS; ++qb) {
            mbar_init(q_full(qb), 2);               // the leader's expect_tx arrive + the peer producer's remote arrive
            mbar_init(q_free(qb), SOFTMAX_WARPS);
        }
        for (int s = 0; s < SLOTS; ++s) {
            mbar_init(kv_full(s), 2);
            mbar_init(kv_free(s), 1);
        }
        for (int b = 0; b < SBUF; ++b) mbar_init(s_full(b), 1);
        for (int b = 0; b < SBUF; ++b)
            for (int c = 0; c < 2; ++c) mbar_init(p_full(b, c), SOFTMAX_WARPS);        // 4 warps of the owning warpgroup x 2 CTAs
        mbar_init(o_full, 1);
        mbar_init(o_free, 2 * SOFTMAX_WARPS);
        mbar_init(pv_done(0), 1);
        mbar_init(pv_done(1), 1);
        fence_barrier_init();
    }
    cluster_sync_all();        // both CTAs are resident before the pair-wide TMEM allocation
    if (warp == MMA_WARP) {
        tmem_alloc_cg2(tmem_slot, 512);
        tmem_relinquish_cg2();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();        // the peer's barriers are initialised before any remote arrive / multicast commit reaches them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == TMA_WARP) {
        if (elect_one()) {
            // ------------------------------ TMA producer (both CTAs) ------------------------------
            auto load_q = [&](int k) {
                const int qb = k % QBUFS;
                if (k >= QBUFS) mbar_wait(q_free(qb), ((k / QBUFS) - 1) & 1, 0x120 + qb);
                const uint32_t lf = mapa_shared(q_full(qb), 0);
                if (rank == 0) mbar_expect_tx(q_full(qb), 2 * Q_BYTES);
                const int head = item_head(k), row0 = item_row0(k);
                tma_load_2d_cg2(q_smem + qb * Q_BYTES, &tmQ, lf, head * HD, row0);
                tma_load_2d_cg2(q_smem + qb * Q_BYTES + QBOX_BYTES, &tmQ, lf, head * HD + 64, row0);
                if (rank != 0) mbar_arrive_cluster(lf);
            };
            load_q(0);
            // load L carries my halves of (K of global step L, V of global step L-2) into ring slot L % SLOTS, one barrier
            int kk = 0, kj = 0;                          // item / step of global step L      (K part)
            int vk = 0, vj = -2;                         // item / step of global step L - 2  (V part)
            for (int L = 0; L <= total + 1; ++L) {
                const bool has_k = L < total, has_v = L >= 2;
                if (has_k && kj == 0 && kk + 1 < n_my) load_q(kk + 1);       // a whole item ahead of its first QK^T
                if (has_k || has_v) {
                    const int slot = L % SLOTS;
                    if (L >= SLOTS) mbar_wait(kv_free(slot), ((L / SLOTS) - 1) & 1, 0x110 + slot);
                    const uint32_t lf = mapa_shared(kv_full(slot), 0);
                    if (rank == 0) mbar_expect_tx(kv_full(slot), 2 * ((has_k ? KHALF_BYTES : 0) + (has_v ? VHALF_BYTES : 0)));
                    const uint32_t dst = kv_smem + slot * SLOT_BYTES;
                    if (has_k) {
                        const int head = item_head(kk);
                        const int row0 = kj * BKV + static_cast<int>(rank) * 64;               // my 64 keys of the tile
                        tma_load_2d_cg2(dst, &tmK, lf, head * HD, row0);
                        tma_load_2d_cg2(dst + KBOX_BYTES, &tmK, lf, head * HD + 64, row0);
                    }
                    if (has_v)                                                                  // all 128 keys, my 64 d columns
                        tma_load_2d_cg2(dst + KHALF_BYTES, &tmV, lf, item_head(vk) * HD + static_cast<int>(rank) * 64, vj * BKV);
                    if (rank != 0) mbar_arrive_cluster(lf);
                }
                if (++kj == n_kv) { kj = 0; ++kk; }
                if (++vj == n_kv) { vj = 0; ++vk; }
            }
        }
    } else if (warp == MMA_WARP) {
        if (rank == 0 && elect_one()) {
            // ------------------------------ MMA issuer (leader CTA) ------------------------------
            auto issue_qk = [&](int sb, uint32_t q_addr, uint32_t k_addr) {
                const uint32_t d = tmem_base + sb * 128;
#pragma unroll
                for (int kk = 0; kk < HD / 16; ++kk) {
                    const uint32_t qoff = (kk >> 2) * QBOX_BYTES + (kk & 3) * 32;
                    const uint32_t koff = (kk >> 2) * KBOX_BYTES + (kk & 3) * 32;
                    umma_ss_cg2(d, make_smem_desc_sw128(q_addr + qoff, 16, 1024), make_smem_desc_sw128(k_addr + koff, 16, 1024),
                                IDESC_QK, kk != 0 ? 1u : 0u);
                }
            };
            // O += P[:, keys of hand-over c] V[keys of hand-over c, :]; hand-over 0 = the first HO0_GROUPS groups of 16 keys
            auto issue_pv = [&](int sb, uint32_t v_addr, bool accumulate, int c) {
                const uint32_t d = tmem_base + O_COL;
                const uint32_t pa = tmem_base + sb * 128;
                const int k0 = c == 0 ? 0 : HO0_GROUPS, k1 = c == 0 ? HO0_GROUPS : BKV / GC;
#pragma unroll
                for (int kk = 0; kk < BKV / GC; ++kk) {
                    if (kk < k0 || kk >= k1) continue;
                    umma_ts_cg2(d, pa + kk * 8, make_smem_desc_sw128(v_addr + kk * 2048, VHALF_BYTES, 1024), IDESC_PV,
                                (accumulate || kk != 0) ? 1u : 0u);
                }
            };
            // QK^T of global step G2 (item k2, step j2): waits for the item's Q at its first step
            int k2 = 0, j2 = 0;
            auto qk_step = [&](int G2, uint32_t k_addr) {
                const int qb = k2 % QBUFS;
                if (j2 == 0) {
                    mbar_wait(q_full(qb), (k2 / QBUFS) & 1, 0x210 + qb);
                    tc_fence_after();
                }
                const int sb = G2 % SBUF;
                issue_qk(sb, q_smem + qb * Q_BYTES, k_addr);
                tc_commit_cg2(s_full(sb), 0x3);
                if (++j2 == n_kv) { j2 = 0; ++k2; }
            };
            for (int G2 = 0; G2 < 2 && G2 < total; ++G2) {               // the first two steps: loads 0, 1 carry K only
                mbar_wait(kv_full(G2), 0, 0x200 + G2);
                tc_fence_after();
                qk_step(G2, kv_smem + G2 * SLOT_BYTES);
                tc_commit_cg2(kv_free(G2), 0x3);
            }
            int b = 0;                                   // G % 3
            uint32_t b_round = 0;                        // G / 3
            int k = 0, j = 0;                            // item / step of global step G
#pragma unroll 1
            for (int G = 0; G < total; ++G) {
                const int L = G + 2, slot = L % SLOTS;
                mbar_wait(kv_full(slot), (L / SLOTS) & 1, 0x200 + slot);              // K of step G+2 and V of step G, one probe
                tc_fence_after();
                const uint32_t base = kv_smem + slot * SLOT_BYTES;
                if (G + 2 < total) qk_step(G + 2, base);     // two steps ahead, into the buffer whose P was consumed by PV(G-1)
                if (j == 0 && k > 0) {
                    // the first PV of an item overwrites O: the epilogue of the item before must have read it out
                    mbar_wait(o_free, (k - 1) & 1, 0x230);
                    tc_fence_after();
                }
                mbar_wait(p_full(b, 0), b_round & 1, 0x220);
                tc_fence_after();
                issue_pv(b, base + KHALF_BYTES, j > 0, 0);
                if (HO0_GROUPS < BKV / GC) {
                    mbar_wait(p_full(b, 1), b_round & 1, 0x221);
                    tc_fence_after();
                    issue_pv(b, base + KHALF_BYTES, true, 1);
                }
                tc_commit_cg2(kv_free(slot), 0x3);
                tc_commit_cg2(pv_done(G & 1), 0x3);
                if (j + 1 == n_kv) tc_commit_cg2(o_full, 0x3);
                if (++b == SBUF) { b = 0; ++b_round; }
                if (++j == n_kv) { j = 0; ++k; }
            }
        }
    } else {
        // ------------------------------ softmax warps (both CTAs) ------------------------------
        // Warpgroup g (warps 4g .. 4g+3) owns the global KV steps G = g (mod 2); one query row per thread (all 128 score
        // columns).  The two threads of a row (same scheduler, warps w and w+4) work on consecutive KV steps half a
        // period apart.  Within an item they share the reference point m of the stored exponentials, handed from the
        // thread of step j-1 to the thread of step j through shared memory; it moves only when the exact row maximum of
        // a tile exceeds it by 2^REF_MARGIN (lazy rescale of O and l).
        const int g = warp >> 2;
        const int quarter = warp & 3;               // TMEM lane quarter accessible to this warp
        const int r = quarter * 32 + lane;
        const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t o_tmem = tmem_base + O_COL + lane_sel;
        const float sl2 = p.scale_log2;
        const uint64_t sl2_2 = f2_pack(sl2, sl2);
        const int tail_valid = p.sk - (n_kv - 1) * BKV;              // valid keys in the last KV tile (1..128)
        const uint32_t bar_mine = 1 + g * 4 + quarter;               // I arrive here once I have published m for my step
        const uint32_t bar_other = 1 + (1 - g) * 4 + quarter;        // ... and wait here for the m of the step before
        const uint32_t bar_pair = 9 + quarter;                       // both threads of the row (epilogue)
        const uint32_t m_addr = xchg + r * 4;
        const uint32_t l_addr = xchg + (BQ + g * BQ + r) * 4, l_other_addr = xchg + (BQ + (1 - g) * BQ + r) * 4;
        const uint32_t o_free_leader = mapa_shared(o_free, 0);

        int b = g;                    // G % 3 of my current step
        uint32_t b_round = 0;         // G / 3
#pragma unroll 1
        for (int k = 0; k < n_my; ++k) {
            const int G0 = k * n_kv;
            float m_last = -INFINITY;     // the reference my l is expressed in
            float l = 0.f;                // sum over MY steps of this item
#pragma unroll 1
            for (int j = (g ^ (G0 & 1)); j < n_kv; j += 2) {
                const int G = G0 + j;
                const uint32_t s_tmem = tmem_base + b * 128 + lane_sel;        // S buffer of step G; P aliases its columns [0,64)
                mbar_wait(s_full(b), b_round & 1, 0x300 + b);
                tc_fence_after();
                uint32_t s[BKV];
                tmem_ld_32x32b_x32(s_tmem + 0, s + 0);
                tmem_ld_32x32b_x32(s_tmem + 32, s + 32);
                tmem_ld_32x32b_x32(s_tmem + 64, s + 64);
                tmem_ld_32x32b_x32(s_tmem + 96, s + 96);
                tc_wait_ld();
                if (j == n_kv - 1 && tail_valid < BKV) {
#pragma unroll
                    for (int c = 0; c < BKV; ++c)
                        if (c >= tail_valid) s[c] = 0xff800000u;   // -inf
                }
                const float mx = row_max<BKV, 0, BKV>(s, -INFINITY);     // exact row maximum of this tile
                float m_prev = -INFINITY;
                if (j > 0) {
                    // The thread of step j-1 (other warpgroup, same scheduler) has published its m: producer / consumer
                    // named barrier (bar.arrive by the publisher, bar.sync here).  It cannot be signalled twice before I
                    // consume it: the publisher's next step needs MY decision first.
                    named_bar_sync(bar_other, 64);
                    m_prev = __uint_as_float(ld_shared_volatile_u32(m_addr));
                }
                const float m_new = ((mx - m_prev) * sl2 > REF_MARGIN) ? mx : m_prev;     // step 0: m_prev = -inf -> mx
                st_shared_u32(m_addr, __float_as_uint(m_new));
                if (j + 1 < n_kv) named_bar_arrive(bar_mine, 64);
                // Rare path (warp-uniform): some row of this warp moves its reference, or my l is in an older reference
                // (m_last <= m_prev <= m_new, so one comparison covers both).
                if (__any_sync(0xffffffffu, m_new != m_last)) {
                    if (j > 0 && __any_sync(0xffffffffu, m_new != m_prev)) {
                        // O holds steps < j relative to m_prev and PV(G-1) may still be accumulating: wait for it, rescale
                        // my row.  Nobody else touches O meanwhile: PV(G) needs my P, and the thread of step j+1 can only
                        // rescale after PV(G).  PV(G-1) is the ((G-1)/2)-th commit on the OTHER parity's barrier.
                        mbar_wait(pv_done(1 - g), ((G - 1) >> 1) & 1, 0x320);
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
                    l *= fast_exp2((m_last - m_new) * sl2);       // 1 if unchanged; first own step: l = 0, exp2(-inf) = 0
                }
                m_last = m_new;
                const float neg_m = -m_new * sl2;
                const uint64_t negm_2 = f2_pack(neg_m, neg_m);
                float lsum = 0.f;
#pragma unroll
                for (int q8 = 0; q8 < BKV / GC; ++q8) {
                    uint32_t pk[GC / 2];
                    switch (q8) {   // compile-time after unrolling
                        case 0: lsum += exp_chunk<BKV, 0 * GC, 1 * GC, 0>(s, pk, sl2_2, negm_2); break;
                        case 1: lsum += exp_chunk<BKV, 1 * GC, 2 * GC, 0>(s, pk, sl2_2, negm_2); break;
                        case 2: lsum += exp_chunk<BKV, 2 * GC, 3 * GC, 0>(s, pk, sl2_2, negm_2); break;
                        case 3: lsum += exp_chunk<BKV, 3 * GC, 4 * GC, 0>(s, pk, sl2_2, negm_2); break;
                        case 4: lsum += exp_chunk<BKV, 4 * GC, 5 * GC, 0>(s, pk, sl2_2, negm_2); break;
                        case 5: lsum += exp_chunk<BKV, 5 * GC, 6 * GC, 0>(s, pk, sl2_2, negm_2); break;
                        case 6: lsum += exp_chunk<BKV, 6 * GC, 7 * GC, 0>(s, pk, sl2_2, negm_2); break;
                        default: lsum += exp_chunk<BKV, 7 * GC, 8 * GC, 0>(s, pk, sl2_2, negm_2); break;
                    }
                    store_p<GC / 2>(s_tmem + q8 * (GC / 2), pk);
                    if (q8 == HO0_GROUPS - 1 || q8 == BKV / GC - 1) {
                        // hand-over: my P columns are in TMEM (wait::st), then one arrival per warp on the LEADER's barrier
                        tc_wait_st();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(mapa_shared(p_full(b, q8 == HO0_GROUPS - 1 ? 0 : 1), 0));
                    }
                }
                l += lsum;
                b += 2;
                if (b >= SBUF) { b -= SBUF; ++b_round; }
            }

            // ------------------------------ epilogue of item k: O / l -> global ------------------------------
            // Meanwhile the tensor core works on QK^T of the next item's first steps (issued two steps ahead).
            named_bar_sync(bar_pair, 64);                                // both threads of the row are past their last step
            const float m_fin = __uint_as_float(ld_shared_volatile_u32(m_addr));
            l *= fast_exp2((m_last - m_fin) * sl2);                      // no own step (n_kv = 1): 0 * exp2(-inf) = 0
            st_shared_u32(l_addr, __float_as_uint(l));
            named_bar_sync(bar_pair, 64);
            const float inv_l = 1.0f / (l + __uint_as_float(ld_shared_volatile_u32(l_other_addr)));
            mbar_wait(o_full, k & 1, 0x310);
            tc_fence_after();
            // Every MMA of the item has completed: its Q buffer is dead and stages the output (4 KB per warp).
            const int qb = k % QBUFS;
            const uint32_t stage = q_smem + qb * Q_BYTES + warp * 4096;
            attn::stage_o_warp(o_tmem + g * (HD / 2), inv_l, stage, lane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(o_free_leader);           // my O columns are out of TMEM
            const int head = item_head(k);
            const int wrow0 = item_row0(k) + quarter * 32;
            attn::flush_o_warp(stage, lane, [&](int rr) -> __nv_bfloat16* {
                const int grow = wrow0 + rr;
                if (grow >= p.sq) return nullptr;
                __nv_bfloat16* base;
                if (p.rows_per_peer > 0) {
                    const int dest = grow / p.rows_per_peer;
                    base = p.out_peer[dest < WVD_MAX_PEERS ? dest : 0] + static_cast<long long>(grow - dest * p.rows_per_peer) * p.ldo;
                } else {
                    base = p.out + static_cast<long long>(grow) * p.ldo;
                }
                return base + head * HD + g * (HD / 2);
            });
            // the staging reads are done: the producer may load the Q tile of item k + 2 over it (async proxy after generic)
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_free(qb));
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();        // neither CTA leaves while the peer's MMAs may still read its operands / arrive on its barriers
    if (warp == MMA_WARP) tmem_dealloc_cg2(tmem_base, 512);
}

}  // namespace attn4

// Launch the persistent cta_group::2 kernel.  Same contract as wvd::attn::launch (attention_sm100.cu), which validates the
// arguments.
int attention_cg2p_launch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                          void* const* out_peers, int world, int64_t rows_per_peer, int64_t ldo, int num_heads,
                          int64_t sq, int64_t sk, float scale, cudaStream_t st) {
    using namespace attn4;
    const int64_t width = (int64_t)num_heads * HD;
    CUtensorMap tmQ, tmK, tmV;
    int rc = get_tensor_map_bf16(&tmQ, q, (uint64_t)sq, (uint64_t)width, (uint64_t)ldq, BQ);
    if (rc) return rc;
    rc = get_tensor_map_bf16(&tmK, k, (uint64_t)sk, (uint64_t)width, (uint64_t)ldk, BKV / 2);      // 64 keys x 64 d boxes
    if (rc) return rc;
    rc = get_tensor_map_bf16(&tmV, v, (uint64_t)sk, (uint64_t)width, (uint64_t)ldv, BKV);          // 128 keys x 64 d boxes
    if (rc) return rc;
    Params p;
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
    // resident CTA pairs of this device, asked once per device (cudaOccupancyMaxActiveClusters)
    static int max_clusters[64] = {0};
    int dev = 0;
    WVD_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) dev = 0;
    if (max_clusters[dev] == 0) {
        WVD_CHECK_CUDA(cudaFuncSetAttribute(attention_cg2p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * 1024);
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = SMEM_BYTES;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2;
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        int n = 0;
        WVD_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, attention_cg2p_kernel, &cfg));
        WVD_REQUIRE(n > 0, "attention_cg2p: no resident CTA pair fits on this device");
        max_clusters[dev] = n;
    }
    const long long q_tiles = (sq + BQ - 1) / BQ;
    p.n_qpairs = (int)((q_tiles + 1) / 2);              // whole CTA pairs; a surplus CTA computes rows >= sq and stores nothing
    const long long items = (long long)p.n_qpairs * num_heads;
    WVD_REQUIRE(items < (1ll << 30), "attention_cg2p: too many work items");
    p.n_items = (int)items;
    const bool persistent = p.n_kv >= MIN_KV_PERSISTENT;
    p.n_clusters = persistent ? (int)(items < max_clusters[dev] ? items : max_clusters[dev]) : (int)items;
    attention_cg2p_kernel<<<dim3(2u * (unsigned)p.n_clusters), NUM_THREADS, SMEM_BYTES, st>>>(tmQ, tmK, tmV, p);
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}

}  // namespace wvd
