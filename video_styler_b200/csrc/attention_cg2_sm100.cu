// K8c: self-attention forward for head_dim 128 on tcgen05 / TMEM (sm_100a), CTA-PAIR MMA variant (cta_group::2).
//
//   out[s, h] = softmax(q_h k_h^T * scale) v_h          non-causal, no mask, no dropout
//   (replaces flash_attention(), diffsynth/models/wan_video_dit.py:28-61, for the ~30k-76k-token self-attention)
//
// Same pipeline as attention_pair_sm100.cu -- one 128-row Q tile per CTA, S triple-buffered in TMEM (S_0 S_1 S_2 O = all
// 512 columns, P(j) aliasing S_(j%3)[0,64)), QK^T(j+2) issued ahead of PV(j), two softmax warpgroups alternating KV tiles
// -- but the two CTAs of a cluster (adjacent Q tiles of one head) are driven by ONE stream of tcgen05.mma cta_group::2
// instructions with M = 256:
//   S  = Q K^T   A = Q from BOTH CTAs' shared memory (128 rows each), B = the K tile SPLIT over the pair by keys
//                (each CTA holds 64 keys x 128 d), D = each CTA's own S rows in its own TMEM
//   O += P V     A = P from both CTAs' TMEM, B = the V tile split over the pair by head-dim columns (each CTA holds
//                128 keys x 64 d -- one TMA box), D = each CTA's own O rows
// Against the cta_group::1 pair kernel, where both CTAs kept a full multicast copy of every K/V tile, each SM stages HALF
// the K/V bytes (32 instead of 64 KB per KV step: half the TMA fill and crossbar traffic, room for a 5-deep ring) and
// reads a third fewer operand bytes from shared memory per MMA (Q 4 KB + K 2 KB instead of 4 + 4; V 2 instead of 4), and
// one thread issues for both tiles (half the issue / barrier work per tile).  The step is power-capped (DESIGN.md
// section 4): bytes moved are clock speed.
// What sets the period (in-kernel timeline, profiles/r2_attention_cg2_timeline.txt): QK^T(j+2) overwrites the buffer of
// P(j-1), so it is issued after the LAST hand-over of softmax(j-1); from there ~2,090 cycles of issue, mbarrier wake-up and
// MMA latency pass until the softmax warps see S(j+2), then ~2,300 cycles of softmax(j+2): 4,390 cycles per THREE KV steps.
// Tried on this kernel and dropped (tools/attn_ab.py, sustained A/B, profiles/r2_attention_ab.txt):
//   exp2 of 1 or 2 of every 4 column pairs on the FMA pipes ........................ +4 % / +12 % time
//   row maximum in the shadow of speculative exponentials against the stale reference  +5 %
//   P in its own TMEM buffers (S_0 S_1 | P_0 P_1 | O) and separate QK^T / PV issuing threads, QK^T(j+2) issued as soon as
//   softmax(j) has LOADED S(j): the softmax warps never wait for S any more, but ... +2 % (an extra mbarrier wait per tile on
//   their path, both warpgroups contending all the time)
//   the next step's K/V wait hoisted above the P waits + local instead of cluster-mapped arrive on the leader ... +3 %
// Only the LEADER CTA (cluster rank 0) issues MMAs.  Both CTAs' TMA loads complete on the leader's barriers
// (cp.async.bulk.tensor .cta_group::2), both CTAs' softmax warps hand P over on the leader's barriers (remote arrive),
// tcgen05.commit is multicast to the S-full / slot-free / PV-done barriers of both CTAs.
//   warps 0-3 / 4-7     softmax warpgroups: one query row per thread, warpgroup g owns the KV tiles j = g (mod 2)
//   warp 8 (1 thread)   TMA producer: my Q tile, then my half of (K_{j+2}, V_j) per step, one ring slot + ONE barrier
//   warp 9 (1 thread)   MMA issuer (leader CTA only); the warp owns the pair-wide TMEM allocation in both CTAs
#include <math.h>

#include "host_utils.h"
#include "ptx.cuh"
#include "softmax_math.cuh"

namespace wvd {
namespace attn3 {

using attn::exp_chunk;
using attn::row_max;
using attn::store_p;

constexpr int BQ = 128, BKV = 128, HD = 128;
constexpr int GC = 16;                        // columns per exp2 / store group
constexpr int HO0_GROUPS = 6;                 // groups of 16 keys in the first hand-over of P (8 = a single hand-over)
constexpr int Q_BYTES = 128 * 128 * 2;        // 32 KB: two 64-column boxes of 128 rows
constexpr int QBOX_BYTES = Q_BYTES / 2;
constexpr int KHALF_BYTES = 64 * 128 * 2;     // my 64 keys x 128 d: two 64-column boxes of 64 rows (8 KB each)
constexpr int KBOX_BYTES = KHALF_BYTES / 2;
constexpr int VHALF_BYTES = 128 * 64 * 2;     // 128 keys x my 64 d columns: one box
constexpr int SLOT_BYTES = KHALF_BYTES + VHALF_BYTES;      // 32 KB: (K_{j+2}, V_j) halves
constexpr int SLOTS = 5;
constexpr int SBUF = 3;                       // S buffers in TMEM
constexpr int O_COL = SBUF * 128;             // first TMEM column of the O accumulator
constexpr int SOFTMAX_WARPS = 8, TMA_WARP = 8, MMA_WARP = 9;
constexpr int NUM_THREADS = 10 * 32;
constexpr int BAR_BYTES = 384;
constexpr int XCHG_BYTES = 3 * BQ * 4;        // m[row], l[warpgroup][row] fp32
constexpr int SMEM_BYTES = Q_BYTES + SLOTS * SLOT_BYTES + BAR_BYTES + XCHG_BYTES + 1024;
constexpr uint32_t IDESC_QK = make_idesc_bf16(256, 128, 0, 0);   // A = Q (K-major), B = K (K-major), M = 256 over the pair
constexpr uint32_t IDESC_PV = make_idesc_bf16(256, 128, 0, 1);   // A = P (TMEM), B = V (MN-major)
constexpr float REF_MARGIN = 8.0f;

struct Params {
    __nv_bfloat16* out;
    long long ldo;
    __nv_bfloat16* out_peer[WVD_MAX_PEERS];   // Ulysses return trip fused into the epilogue (see attention_sm100.cu)
    int rows_per_peer;
    int sq, sk, n_kv;
    float scale_log2;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
attention_cg2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;     // same offset in both CTAs of the pair
    uint8_t* smem_gen = smem_raw + (smem_base - raw_u32);
    const uint32_t q_smem = smem_base;
    const uint32_t kv_smem = smem_base + Q_BYTES;
    const uint32_t bar_base = kv_smem + SLOTS * SLOT_BYTES;
    const uint32_t q_full = bar_base;                                            // LEADER: both CTAs' Q tiles have landed
    auto kv_full = [&](int s) { return bar_base + 8 + s * 8; };                 // LEADER: both halves of slot s have landed
    auto kv_free = [&](int s) { return bar_base + 64 + s * 8; };                // the MMAs reading slot s have completed
    auto s_full = [&](int b) { return bar_base + 128 + b * 8; };                // S buffer b holds Q K^T
    // LEADER: hand-over c of P of the tile in S buffer b is in TMEM of BOTH CTAs.  Per BUFFER, not per warpgroup: with S
    // triple-buffered the softmax warps can hand over tiles j and j+2 before the MMA issuer has consumed tile j (two
    // phases of a per-warpgroup barrier -> parity aliasing -> deadlock); tile j+3 cannot be handed over before PV(j).
    auto p_full = [&](int b, int c) { return bar_base + 160 + (b * 2 + c) * 8; };
    auto pv_done = [&](int g) { return bar_base + 208 + g * 8; };              // PV of warpgroup g's latest tile (and every PV before it) has completed
    const uint32_t o_full = bar_base + 224;
    const uint32_t tmem_slot = bar_base + 240;
    const uint32_t xchg = bar_base + BAR_BYTES;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Q_BYTES + SLOTS * SLOT_BYTES + 240);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int head = blockIdx.y;
    const int q_row0 = blockIdx.x * BQ;
    const int n_kv = p.n_kv;
    const uint32_t rank = cluster_ctarank();

    if (warp == TMA_WARP && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == MMA_WARP && lane == 0) {
        mbar_init(q_full, 2);                       // the leader's expect_tx arrive + the peer producer's remote arrive
        for (int s = 0; s < SLOTS; ++s) {
            mbar_init(kv_full(s), 2);
            mbar_init(kv_free(s), 1);
        }
        for (int b = 0; b < SBUF; ++b) mbar_init(s_full(b), 1);
        for (int b = 0; b < SBUF; ++b)
            for (int c = 0; c < 2; ++c) mbar_init(p_full(b, c), SOFTMAX_WARPS);        // 4 warps of the owning warpgroup x 2 CTAs
        mbar_init(o_full, 1);
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
            {
                const uint32_t lf = mapa_shared(q_full, 0);
                if (rank == 0) mbar_expect_tx(q_full, 2 * Q_BYTES);
                tma_load_2d_cg2(q_smem, &tmQ, lf, head * HD, q_row0);
                tma_load_2d_cg2(q_smem + QBOX_BYTES, &tmQ, lf, head * HD + 64, q_row0);
                if (rank != 0) mbar_arrive_cluster(lf);
            }
            // load q = j + 2 carries my halves of (K_q, V_{q-2}) into ring slot q % SLOTS, one barrier for both
            for (int q = 0; q <= n_kv + 1; ++q) {
                const bool has_k = q < n_kv, has_v = q >= 2;
                if (!has_k && !has_v) continue;                          // q = 1 of a single-tile sequence
                const int slot = q % SLOTS;
                if (q >= SLOTS) mbar_wait(kv_free(slot), ((q / SLOTS) - 1) & 1, 0x110 + slot);
                const uint32_t lf = mapa_shared(kv_full(slot), 0);
                if (rank == 0) mbar_expect_tx(kv_full(slot), 2 * ((has_k ? KHALF_BYTES : 0) + (has_v ? VHALF_BYTES : 0)));
                const uint32_t dst = kv_smem + slot * SLOT_BYTES;
                if (has_k) {
                    const int row0 = q * BKV + static_cast<int>(rank) * 64;                 // my 64 keys of the tile
                    tma_load_2d_cg2(dst, &tmK, lf, head * HD, row0);
                    tma_load_2d_cg2(dst + KBOX_BYTES, &tmK, lf, head * HD + 64, row0);
                }
                if (has_v)                                                                  // all 128 keys, my 64 d columns
                    tma_load_2d_cg2(dst + KHALF_BYTES, &tmV, lf, head * HD + static_cast<int>(rank) * 64, (q - 2) * BKV);
                if (rank != 0) mbar_arrive_cluster(lf);
            }
        }
    } else if (warp == MMA_WARP) {
        if (rank == 0 && elect_one()) {
            // ------------------------------ MMA issuer (leader CTA) ------------------------------
            auto issue_qk = [&](int sb, uint32_t k_addr) {
                const uint32_t d = tmem_base + sb * 128;
#pragma unroll
                for (int kk = 0; kk < HD / 16; ++kk) {
                    const uint32_t qoff = (kk >> 2) * QBOX_BYTES + (kk & 3) * 32;
                    const uint32_t koff = (kk >> 2) * KBOX_BYTES + (kk & 3) * 32;
                    umma_ss_cg2(d, make_smem_desc_sw128(q_smem + qoff, 16, 1024), make_smem_desc_sw128(k_addr + koff, 16, 1024),
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
            mbar_wait(q_full, 0, 0x210);
            tc_fence_after();
            for (int q = 0; q < 2 && q < n_kv; ++q) {                    // K_0, K_1 (slots 0, 1; S buffers 0, 1)
                mbar_wait(kv_full(q), 0, 0x200 + q);
                tc_fence_after();
                issue_qk(q, kv_smem + q * SLOT_BYTES);
                tc_commit_cg2(s_full(q), 0x3);
                tc_commit_cg2(kv_free(q), 0x3);
            }
            int b = 0;                                   // j % 3
            uint32_t b_round = 0;                        // j / 3
#pragma unroll 1
            for (int j = 0; j < n_kv; ++j) {
                const int q = j + 2, slot = q % SLOTS, sb2 = b == 0 ? 2 : b - 1;      // sb2 = (j + 2) % 3
                mbar_wait(kv_full(slot), (q / SLOTS) & 1, 0x200 + slot);              // K_{j+2} and V_j, one probe
                tc_fence_after();
                const uint32_t base = kv_smem + slot * SLOT_BYTES;
                if (j + 2 < n_kv) {
                    // QK^T two steps ahead, into the buffer whose P was consumed by PV(j-1) (issued in the last iteration)
                    issue_qk(sb2, base);
                    tc_commit_cg2(s_full(sb2), 0x3);
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
                tc_commit_cg2(pv_done(j & 1), 0x3);
                if (j + 1 == n_kv) tc_commit_cg2(o_full, 0x3);
                if (++b == SBUF) { b = 0; ++b_round; }
            }
        }
    } else {
        // ------------------------------ softmax warps (both CTAs) ------------------------------
        // Warpgroup g (warps 4g .. 4g+3) owns the KV tiles j = g (mod 2); one query row per thread (all 128 score
        // columns).  The two threads of a row (same scheduler, warps w and w+4) work on consecutive KV tiles half a
        // period apart, so together they keep the MUFU pipe fed although a lone warp cannot, and neither ever waits for
        // the tensor core: S(j+2) is produced during tile j.  The only thing they share is the reference point m of the
        // stored exponentials, handed from the thread of tile j-1 to the thread of tile j through shared memory: it
        // moves only when the exact row maximum of a tile exceeds it by 2^REF_MARGIN (lazy rescale of O and l).
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
        float m_last = -INFINITY;     // the reference my l is expressed in
        float l = 0.f;                // sum over MY tiles

        int b = g;                    // j % 3 of my current tile
        uint32_t b_round = 0;         // j / 3
#pragma unroll 1
        for (int j = g; j < n_kv; j += 2) {
            const uint32_t s_tmem = tmem_base + b * 128 + lane_sel;        // S buffer of tile j; P aliases its columns [0,64)
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
                // The thread of tile j-1 (other warpgroup, same scheduler) has published its m: producer / consumer
                // named barrier (bar.arrive by the publisher, bar.sync here).  It cannot be signalled twice before I
                // consume it: the publisher's next tile needs MY decision first.
                named_bar_sync(bar_other, 64);
                m_prev = __uint_as_float(ld_shared_volatile_u32(m_addr));
            }
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

        // ------------------------------ epilogue: O / l -> global ------------------------------
        named_bar_sync(bar_pair, 64);                                // both threads of the row are past their last tile
        const float m_fin = __uint_as_float(ld_shared_volatile_u32(m_addr));
        l *= fast_exp2((m_last - m_fin) * sl2);                      // no own tile (n_kv = 1, g = 1): 0 * exp2(-inf) = 0
        st_shared_u32(l_addr, __float_as_uint(l));
        named_bar_sync(bar_pair, 64);
        const float inv_l = 1.0f / (l + __uint_as_float(ld_shared_volatile_u32(l_other_addr)));
        mbar_wait(o_full, 0, 0x310);
        tc_fence_after();
        // All MMAs of the pair have completed (o_full): my CTA's Q tile is dead, its 32 KB stage the output (4 KB per warp).
        const int wrow0 = q_row0 + quarter * 32;
        attn::store_o_warp_coalesced(o_tmem + g * (HD / 2), inv_l, q_smem + warp * 4096, lane, [&](int rr) -> __nv_bfloat16* {
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
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();        // neither CTA leaves while the peer's MMAs may still read its operands / arrive on its barriers
    if (warp == MMA_WARP) tmem_dealloc_cg2(tmem_base, 512);
}

}  // namespace attn3

int attn_cg2_read_diag(unsigned long long* out) {
    if (cudaMemcpyFromSymbol(out, g_diag, sizeof(unsigned long long) * 8) != cudaSuccess) return -1;
    unsigned long long zero[8] = {0};
    cudaMemcpyToSymbol(g_diag, zero, sizeof(zero));
    return 0;
}

// Launch the cta_group::2 kernel.  Same contract as wvd::attn::launch (attention_sm100.cu), which validates the arguments.
int attention_cg2_launch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                         void* const* out_peers, int world, int64_t rows_per_peer, int64_t ldo, int num_heads,
                         int64_t sq, int64_t sk, float scale, cudaStream_t st) {
    using namespace attn3;
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
    static unsigned long long configured = 0;
    if (first_use_on_current_device(&configured))
        WVD_CHECK_CUDA(cudaFuncSetAttribute(attention_cg2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    unsigned q_tiles = (unsigned)((sq + BQ - 1) / BQ);
    q_tiles = (q_tiles + 1u) & ~1u;          // whole CTA pairs; a surplus CTA computes rows >= sq and stores nothing
    dim3 grid(q_tiles, (unsigned)num_heads);
    attention_cg2_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmQ, tmK, tmV, p);
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
}

}  // namespace wvd
