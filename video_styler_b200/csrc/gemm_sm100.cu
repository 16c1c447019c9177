// K5-K7/K10: C = epilogue(A . W^T + bias) on tcgen05 tensor cores (sm_100a).
//
// Persistent, warp-specialised kernel in two variants of one template (CG = CTAs per MMA):
//   CG = 2  one 256 x 256 output tile per CTA PAIR (2-CTA cluster): tcgen05.mma cta_group::2, M = 256 across the pair.
//           Each CTA stages ITS 128 rows of A and ITS 128 rows of W per k-block (32 KB per stage, 6 stages) -- half the
//           shared-memory operand traffic and half the TMA write traffic per SM of the single-CTA form, which is what
//           bounds a 128 x 256 x 16 MMA stream on one SM (operand reads 96 B/clk + TMA fill 96 B/clk against a
//           128 B/clk port).  Only the leader CTA issues MMAs; both CTAs' TMA loads count on the LEADER's full barrier;
//           tcgen05.commit is multicast to both CTAs' empty / accumulator-full barriers; both CTAs' epilogue warps
//           release the accumulator stage on the leader's barrier (remote arrive).
//   CG = 1  one 128 x 256 tile per CTA (48 KB per stage, 4 stages): kept for shapes whose tile count quantises badly
//           over 74 CTA pairs (e.g. M = 3,705 rows per rank at 8 GPUs: 300 pair tiles = 4.05 waves) -- the dispatcher
//           picks the variant with the smaller modelled time.
//   warp 0 (1 thread)  TMA producer: A (128 x 64) and W (256/CG x 64) bf16 tiles, 128-byte swizzle, mbarrier ring
//   warp 1 (1 thread)  MMA issuer (leader CTA only for CG = 2): 128*CG x 256 x 16, fp32 accumulators in TMEM
//   warp 2             TMEM allocator (512 columns = 2 accumulator stages of 256 columns)
//   warps 4-7          epilogue: tcgen05.ld (one accumulator row per thread) -> bias / GELU-tanh / gate / residual in
//                      fp32 with the reference's bf16 rounding points -> 128-byte-swizzled staging tile in shared memory
//                      -> TMA store (32 rows x 64 columns per warp, double-buffered, full 128-byte lines; the M / N tails
//                      are clipped by the TMA unit)
// The epilogue of tile i overlaps the main loop of tile i+1 through the two TMEM accumulator stages.
// Tiles are rasterised in bands of 8 N-tiles so that the W band (<= 2048 x K) stays L2-resident while A streams.
// GROUPED launch: up to 3 weight matrices of the same shape share A and one launch (q | k | v projections): the tile's
// N index selects the weight tensor map / bias, outputs go to consecutive column slices of C.
//
// Replaces the F.linear call sites of the reference (diffsynth/models/wan_video_dit.py:131-134,157-160,209-210;
// wan_video_vace.py:15,21) and fuses GateModule (:189-194), the ungated residual (:227) and nn.GELU('tanh').
#include "host_utils.h"
#include "ptx.cuh"

namespace wvd {
namespace gemm {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int NUM_THREADS = 256;
constexpr int MAX_GROUPS = 3;
constexpr int CHUNK = 64;              // output columns per staging tile (64 bf16 = one 128-byte swizzle row)
constexpr int STORE_TILE_BYTES = 32 * CHUNK * 2;              // 4 KB: 32 rows x 128 B
constexpr int STORE_BYTES = 4 * 2 * STORE_TILE_BYTES;         // 4 epilogue warps x 2 buffers

// CG = CTAs per MMA (1 or 2); MT = 128-row A tiles per CTA (2 only with CG = 2: a 512 x 256 tile per CTA pair whose two
// 256 x 256 halves occupy ALL 512 TMEM columns)
template <int CG, int MT> struct Cfg {
    static_assert(MT == 1 || (MT == 2 && CG == 2), "two A tiles per CTA only in the CTA-pair form");
    static constexpr int B_ROWS = BN / CG;                    // rows of W this CTA stages per k-block
    static constexpr int A_TILE_BYTES = BM * BK * 2;          // 16 KB per 128-row A tile
    static constexpr int A_BYTES = MT * A_TILE_BYTES;
    static constexpr int B_BYTES = B_ROWS * BK * 2;           // 32 KB (CG = 1) / 16 KB (CG = 2)
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;     // 48 / 32 / 48 KB
    static constexpr int STAGES = (CG == 2 && MT == 1) ? 6 : 4;
    static constexpr int ROWS = BM * CG * MT;                 // rows of one unit (per CTA pair for CG = 2)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STORE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr uint32_t IDESC = make_idesc_bf16(BM * CG, BN, 0, 0);
};

struct Params {
    int M, N, K;                       // N = columns per group
    int groups, tiles_per_group;
    int m_units, n_tiles, num_units, k_blocks;       // a unit = (128*CG rows) x (256 columns)
    int band;                          // N-tiles per rasterisation band (the W band that stays L2-resident while A streams)
    const __nv_bfloat16* bias[MAX_GROUPS];
    const __nv_bfloat16* gate;
    const __nv_bfloat16* res;
    long long ldr;
};

// Tile schedule: the N-tiles are cut into bands; a band is swept M-major (all M units of a band before the next band),
// so the band's W rows are re-read from L2 while A streams.  Every cluster walks this one common order with a static
// stride.  (Tried and dropped: two halves of the clusters sweeping alternate bands concurrently, to keep each W band on
// one die of the two-die B200 -- 3 % slower sustained, profiles/r2_gemm_sustained.txt.)
__device__ __forceinline__ void unit_coords(const Params& p, int unit, int& m_unit, int& n_blk) {
    const int per_band = p.m_units * p.band;
    const int band = unit / per_band;
    const int within = unit - band * per_band;
    const int n_start = band * p.band;
    const int bw = min(p.band, p.n_tiles - n_start);
    m_unit = within / bw;
    n_blk = n_start + within % bw;
}

__device__ __forceinline__ float gelu_tanh(float x) {
    // 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))), tanh(u) = 1 - 2 / (exp(2u) + 1)
    const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
    const float e = __expf(2.0f * u);
    const float t = 1.0f - __fdividef(2.0f, e + 1.0f);
    return 0.5f * x * (1.0f + t);
}

// Epilogue of 8 consecutive columns of one accumulator row, with the reference's rounding points: F.linear rounds
// acc + bias to bf16 once; GELU-tanh is evaluated in fp32 on that bf16 value and rounded once; GateModule's
// x + gate * y (wan_video_dit.py:189-194) and the plain residual add round after each operation.  The bf16 operations
// run as packed bf16x2 instructions (__hmul2_rn / __hadd2_rn: one rounding each, never contracted), two elements per
// instruction, instead of emulating every rounding in fp32.
template <int EPI>
__device__ __forceinline__ uint4 epilogue_pack8(const Params& p, const __nv_bfloat16* bias, const uint32_t* acc,
                                                long long row, int col, bool row_ok) {
    // acc: 8 fp32 accumulator values (as bits) for columns col..col+7 (within the group) of `row`
    __nv_bfloat162 y[4];
    if (bias != nullptr) {
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(bias + col));
        const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 bf = unpack_bf16x2(bw[q]);
            y[q] = __floats2bfloat162_rn(__uint_as_float(acc[2 * q]) + bf.x, __uint_as_float(acc[2 * q + 1]) + bf.y);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) y[q] = __floats2bfloat162_rn(__uint_as_float(acc[2 * q]), __uint_as_float(acc[2 * q + 1]));
    }
    if (EPI == WVD_EPI_BIAS_GELU) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 yf = __bfloat1622float2(y[q]);
            y[q] = __floats2bfloat162_rn(gelu_tanh(yf.x), gelu_tanh(yf.y));
        }
    }
    if (EPI == WVD_EPI_BIAS_GELU_T5) {
        // the umT5 encoder's own GELU module: every torch op rounds to bf16 (pow, *0.044715, +x, *sqrt(2/pi), tanh, 1+, 0.5*x, *)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 yf = __bfloat1622float2(y[q]);
            float o[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float x = e == 0 ? yf.x : yf.y;
                const float x3 = bf16_round(x * x * x);
                const float t1 = bf16_round(0.044715f * x3);
                const float t2 = bf16_round(x + t1);
                const float t3 = bf16_round(0.7978845608028654f * t2);
                const float t4 = bf16_round(tanhf(t3));
                const float t5 = bf16_round(1.0f + t4);
                const float t6 = bf16_round(0.5f * x);
                o[e] = t6 * t5;
            }
            y[q] = __floats2bfloat162_rn(o[0], o[1]);
        }
    }
    if (EPI == WVD_EPI_BIAS_GATE_RES) {
        const uint4 g = __ldg(reinterpret_cast<const uint4*>(p.gate + col));
        const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) y[q] = __hmul2_rn(*reinterpret_cast<const __nv_bfloat162*>(&gw[q]), y[q]);
    }
    if (EPI == WVD_EPI_BIAS_RES || EPI == WVD_EPI_BIAS_GATE_RES || EPI == WVD_EPI_BIAS_MUL) {
        uint4 r = make_uint4(0u, 0u, 0u, 0u);
        if (row_ok) r = *reinterpret_cast<const uint4*>(p.res + row * p.ldr + col);
        const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
            y[q] = EPI == WVD_EPI_BIAS_MUL ? __hmul2_rn(y[q], *reinterpret_cast<const __nv_bfloat162*>(&rw[q]))
                                           : __hadd2_rn(*reinterpret_cast<const __nv_bfloat162*>(&rw[q]), y[q]);
    }
    uint4 o;
    o.x = *reinterpret_cast<const uint32_t*>(&y[0]);
    o.y = *reinterpret_cast<const uint32_t*>(&y[1]);
    o.z = *reinterpret_cast<const uint32_t*>(&y[2]);
    o.w = *reinterpret_cast<const uint32_t*>(&y[3]);
    return o;
}

template <int EPI, int CG, int MT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
                 const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2,
                 const __grid_constant__ CUtensorMap tmC, const Params p) {
    using C = Cfg<CG, MT>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;       // same offset in both CTAs of a pair
    uint8_t* smem_gen = smem_raw + (smem_base - raw_u32);
    const uint32_t store_base = smem_base + C::STAGES * C::STAGE_BYTES;      // 1024-byte aligned (stage sizes are)
    const uint32_t bar_base = store_base + STORE_BYTES;
    auto full_bar = [&](int s) { return bar_base + s * 8; };
    auto empty_bar = [&](int s) { return bar_base + 64 + s * 8; };
    auto tfull_bar = [&](int a) { return bar_base + 128 + a * 8; };
    auto tempty_bar = [&](int a) { return bar_base + 144 + a * 8; };
    const uint32_t tmem_slot = bar_base + 160;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + C::STAGES * C::STAGE_BYTES + STORE_BYTES + 160);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    const int cluster_id = blockIdx.x / CG;
    const int num_clusters = gridDim.x / CG;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB0);
        tma_prefetch_desc(&tmC);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full_bar(s), CG);          // CG = 2: the leader's expect_tx arrive + the peer producer's remote arrive
            mbar_init(empty_bar(s), 1);          // tcgen05.commit (multicast to both CTAs for CG = 2)
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4 * CG);    // every epilogue warp of every CTA of the pair (on the leader's barrier)
        }
        fence_barrier_init();
    }
    if (CG == 2) cluster_sync_all();             // both CTAs are resident before the pair-wide TMEM allocation
    if (warp == 2) {
        if (CG == 2) { tmem_alloc_cg2(tmem_slot, 512); tmem_relinquish_cg2(); }
        else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();             // the peer's barriers are initialised before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------ TMA producer (both CTAs) ------------------------------
            int stage = 0;
            uint32_t phase = 0;
            for (int unit = cluster_id; unit < p.num_units; unit += num_clusters) {
                int m_unit, n_blk;
                unit_coords(p, unit, m_unit, n_blk);
                const int grp = n_blk / p.tiles_per_group;
                const int n_in = n_blk - grp * p.tiles_per_group;
                const CUtensorMap* tmB = grp == 0 ? &tmB0 : (grp == 1 ? &tmB1 : &tmB2);
                const int a_row = (m_unit * CG + static_cast<int>(rank)) * (BM * MT);       // MT = 2: one 256-row box
                const int b_row = n_in * BN + static_cast<int>(rank) * C::B_ROWS;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1, 0x100 + stage);
                    const uint32_t a_dst = smem_base + stage * C::STAGE_BYTES;
                    if (CG == 1) {
                        mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
                        tma_load_2d(a_dst, &tmA, full_bar(stage), kb * BK, a_row);
                        tma_load_2d(a_dst + C::A_BYTES, tmB, full_bar(stage), kb * BK, b_row);
                    } else {
                        const uint32_t leader_full = mapa_shared(full_bar(stage), 0);
                        if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);     // both CTAs' bytes
                        tma_load_2d_cg2(a_dst, &tmA, leader_full, kb * BK, a_row);
                        tma_load_2d_cg2(a_dst + C::A_BYTES, tmB, leader_full, kb * BK, b_row);
                        if (rank != 0) mbar_arrive_cluster(leader_full);
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // ------------------------------ MMA issuer (leader CTA) ------------------------------
            int stage = 0;
            uint32_t phase = 0;
            auto advance = [](int& st, uint32_t& ph) { if (++st == C::STAGES) { st = 0; ph ^= 1; } };
            // the four k-steps of one k-block into accumulator columns d_tmem, A tile `at` of the stage
            auto issue_kblock = [&](uint32_t d_tmem, int st, int at, bool first_kb) {
                const uint32_t a_addr = smem_base + st * C::STAGE_BYTES + at * C::A_TILE_BYTES;
                const uint32_t b_addr = smem_base + st * C::STAGE_BYTES + C::A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                    const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                    const uint32_t accum = (first_kb && k == 0) ? 0u : 1u;
                    if (CG == 2) umma_ss_cg2(d_tmem, da, db, C::IDESC, accum);
                    else umma_ss(d_tmem, da, db, C::IDESC, accum);
                }
            };
            auto commit = [&](uint32_t bar) { if (CG == 2) tc_commit_cg2(bar, 0x3); else tc_commit(bar); };
            if constexpr (MT == 1) {
                int acc = 0;
                uint32_t acc_phase = 0;
                for (int unit = cluster_id; unit < p.num_units; unit += num_clusters) {
                    mbar_wait(tempty_bar(acc), acc_phase ^ 1, 0x200 + acc);
                    tc_fence_after();
                    for (int kb = 0; kb < p.k_blocks; ++kb) {
                        mbar_wait(full_bar(stage), phase, 0x300 + stage);
                        tc_fence_after();
                        issue_kblock(tmem_base + acc * BN, stage, 0, kb == 0);
                        commit(empty_bar(stage));
                        advance(stage, phase);
                    }
                    commit(tfull_bar(acc));
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                }
            } else {
                // One tile = two 256 x 256 halves (A tile 0 / 1 of each CTA) in TMEM columns [0,256) / [256,512): no
                // second accumulator stage.  The epilogue drains half 0 first; the next tile's first HEAD k-blocks are
                // issued for half 0 ALONE while half 1 is still draining, then half 1 catches up on the same stages.
                uint32_t t_phase = 0;
                for (int unit = cluster_id; unit < p.num_units; unit += num_clusters) {
                    const int head = p.k_blocks < C::STAGES - 1 ? p.k_blocks : C::STAGES - 1;
                    mbar_wait(tempty_bar(0), t_phase ^ 1, 0x200);
                    tc_fence_after();
                    int st = stage;
                    uint32_t ph = phase;
                    for (int kb = 0; kb < head; ++kb) {
                        mbar_wait(full_bar(st), ph, 0x300 + st);
                        tc_fence_after();
                        issue_kblock(tmem_base, st, 0, kb == 0);
                        advance(st, ph);
                    }
                    mbar_wait(tempty_bar(1), t_phase ^ 1, 0x201);
                    tc_fence_after();
                    for (int kb = 0; kb < head; ++kb) {
                        issue_kblock(tmem_base + BN, stage, 1, kb == 0);
                        commit(empty_bar(stage));
                        advance(stage, phase);
                    }
                    for (int kb = head; kb < p.k_blocks; ++kb) {
                        mbar_wait(full_bar(stage), phase, 0x300 + stage);
                        tc_fence_after();
                        issue_kblock(tmem_base, stage, 0, false);
                        issue_kblock(tmem_base + BN, stage, 1, false);
                        commit(empty_bar(stage));
                        advance(stage, phase);
                    }
                    commit(tfull_bar(0));
                    t_phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------ epilogue (both CTAs) ------------------------------
        const int q = warp - 4;   // == warp % 4: the TMEM lane quarter this warp may access
        const uint32_t stage_buf = store_base + q * 2 * STORE_TILE_BYTES;
        const uint32_t my_row_off = static_cast<uint32_t>(lane) * 128u;
        const uint32_t sw = static_cast<uint32_t>(lane & 7);
        int acc = 0;                  // MT = 1: accumulator stage of the current tile; MT = 2: unused (one tile fills TMEM)
        uint32_t acc_phase = 0;
        uint32_t n_stores = 0;        // staging tiles this warp has handed to the TMA unit
        for (int unit = cluster_id; unit < p.num_units; unit += num_clusters) {
            int m_unit, n_blk;
            unit_coords(p, unit, m_unit, n_blk);
            const int grp = n_blk / p.tiles_per_group;
            const int n_in = n_blk - grp * p.tiles_per_group;
            const __nv_bfloat16* bias = p.bias[grp];
            mbar_wait(tfull_bar(MT == 1 ? acc : 0), acc_phase, 0x400 + acc);
            tc_fence_after();
#pragma unroll 1
            for (int at = 0; at < MT; ++at) {
                const int half = MT == 1 ? acc : at;                 // which 256 TMEM columns
                const int row0 = (m_unit * CG + static_cast<int>(rank)) * (BM * MT) + at * BM + q * 32;   // first row of this warp's slab
                const long long row = static_cast<long long>(row0) + lane;
                const uint32_t taddr = tmem_base + half * BN + (static_cast<uint32_t>(q * 32) << 16);
                const bool row_ok = row < p.M;
#pragma unroll 1
                for (int c = 0; c < BN / CHUNK; ++c) {
                    uint32_t v[CHUNK];
                    tmem_ld_32x32b_x32(taddr + c * CHUNK, v);
                    tmem_ld_32x32b_x32(taddr + c * CHUNK + 32, v + 32);
                    tc_wait_ld();
                    if (c == BN / CHUNK - 1) {
                        // these 256 columns are in registers: hand them back to the MMA issuer before the math and stores
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CG == 2) mbar_arrive_cluster(mapa_shared(tempty_bar(half), 0));
                            else mbar_arrive(tempty_bar(half));
                        }
                    }
                    const int col_in = n_in * BN + c * CHUNK;            // column within the group
                    if (col_in >= p.N || row0 >= p.M) continue;         // warp-uniform: nothing of this chunk is stored
                    // the buffer's previous TMA store (two stores ago) must have finished READING shared memory
                    const uint32_t buf = stage_buf + (n_stores & 1u) * STORE_TILE_BYTES;
                    if (lane == 0) tma_store_wait_read<1>();
                    __syncwarp();
#pragma unroll
                    for (int g = 0; g < CHUNK / 8; ++g) {
                        const int col = col_in + g * 8;
                        uint4 o = make_uint4(0u, 0u, 0u, 0u);
                        if (col < p.N) o = epilogue_pack8<EPI>(p, bias, v + g * 8, row, col, row_ok);
                        st_shared_v4(buf + my_row_off + ((static_cast<uint32_t>(g) ^ sw) << 4), o.x, o.y, o.z, o.w);
                    }
                    fence_proxy_async();          // my generic-proxy writes -> visible to the TMA (async proxy)
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmC, buf, grp * p.N + col_in, row0);     // rows >= M / columns >= width are clipped
                        tma_store_commit();
                    }
                    ++n_stores;
                }
            }
            if (MT == 1) {
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            } else {
                acc_phase ^= 1;
            }
        }
        if (lane == 0) tma_store_wait<0>();      // all my stores have completed before the CTA's shared memory goes away
    }

    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();      // neither CTA leaves while the peer may still arrive on its barriers / read its operands
    if (warp == 2) {
        if (CG == 2) tmem_dealloc_cg2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
}

template <int EPI, int CG, int MT>
int launch(const CUtensorMap& tmA, const CUtensorMap* tmB, const CUtensorMap& tmC, const Params& p, cudaStream_t stream) {
    using C = Cfg<CG, MT>;
    static unsigned long long configured = 0;
    if (first_use_on_current_device(&configured))
        WVD_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<EPI, CG, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    const int clusters = p.num_units < sm_count() / CG ? p.num_units : sm_count() / CG;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(clusters * CG));
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    WVD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<EPI, CG, MT>, tmA, tmB[0], tmB[1], tmB[2], tmC, p));
    return WVD_OK;
}

// Variant choice: whole waves of units over the machine, weighted by the measured relative speed of the variants on
// large problems (sustained, power-capped: tools/gemm_sustained.py, profiles/r2_gemm_sustained.txt).  Bigger tiles move
// fewer bytes per flop through L2 / the crossbar / shared memory -- under the power cap that is clock speed -- but
// quantise worse on small M.
static int pick_variant(int64_t M, int64_t n_tiles, int64_t K) {
    const int sms = sm_count();
    auto waves = [&](int rows, int ctas) { return static_cast<double>((((M + rows - 1) / rows) * n_tiles + sms / ctas - 1) / (sms / ctas)); };
    // measured sustained speed of a full machine relative to the 1-CTA variant (29,640-row problems): the 256 x 256 pair
    // tile wins at K = 5,120 (its two accumulator stages hide the whole epilogue), the 512 x 256 pair tile at K = 13,824
    // (fewest operand bytes per flop; its partly exposed epilogue is amortised over 216 k-blocks)
    const bool long_k = K >= 8192;
    const double t1 = waves(BM, 1);                                        // 128 x 256 per CTA
    const double t2 = waves(2 * BM, 2) / (long_k ? 1.05 : 1.15);           // 256 x 256 per pair
    const double t3 = waves(4 * BM, 2) * 2.0 / (long_k ? 1.15 : 1.10);     // 512 x 256 per pair (twice the work per unit)
    if (t3 < t2 && t3 < t1) return WVD_GEMM_2CTA_M512;
    return t2 <= t1 ? WVD_GEMM_2CTA : WVD_GEMM_1CTA;
}

static int run(const void* A, int64_t lda, const void* const* W, int64_t ldw, const void* const* bias, void* Cp,
               int64_t ldc, int64_t M, int64_t N, int64_t K, int groups, int epilogue, const void* gate,
               const void* residual, int64_t ldr, int variant, wvd_stream_t stream, const char* who) {
    WVD_REQUIRE(A && W && Cp, "%s: null pointer", who);
    WVD_REQUIRE(groups >= 1 && groups <= MAX_GROUPS, "%s: 1..%d weight groups (got %d)", who, MAX_GROUPS, groups);
    WVD_REQUIRE(M > 0 && N > 0 && K > 0, "%s: empty problem M=%lld N=%lld K=%lld", who, (long long)M, (long long)N, (long long)K);
    WVD_REQUIRE(M < (1ll << 31) && N * groups < (1ll << 31) && K < (1ll << 31), "%s: dimension too large", who);
    WVD_REQUIRE(K % 8 == 0 && N % 8 == 0, "%s: K and N must be multiples of 8 (got K=%lld N=%lld)", who, (long long)K, (long long)N);
    WVD_REQUIRE(groups == 1 || N % CHUNK == 0, "%s: grouped launches need N %% 64 == 0 (got %lld)", who, (long long)N);
    WVD_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0 && lda >= K && ldw >= K && ldc >= N * groups,
                "%s: leading dims must be multiples of 8 and cover the row", who);
    WVD_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)Cp % 16 == 0) && ((uintptr_t)gate % 16 == 0) && ((uintptr_t)residual % 16 == 0),
                "%s: pointers must be 16-byte aligned", who);
    for (int g = 0; g < groups; ++g)
        WVD_REQUIRE(W[g] && ((uintptr_t)W[g] % 16 == 0) && (bias == nullptr || (uintptr_t)bias[g] % 16 == 0),
                    "%s: weight / bias pointers must be non-null (weights) and 16-byte aligned", who);
    WVD_REQUIRE(epilogue >= WVD_EPI_BIAS && epilogue <= WVD_EPI_BIAS_GELU_T5, "%s: bad epilogue %d", who, epilogue);
    WVD_REQUIRE(groups == 1 || epilogue == WVD_EPI_BIAS, "%s: grouped launches take the plain bias epilogue", who);
    if (epilogue == WVD_EPI_BIAS_RES || epilogue == WVD_EPI_BIAS_GATE_RES || epilogue == WVD_EPI_BIAS_MUL)
        WVD_REQUIRE(residual && ldr % 8 == 0 && ldr >= N, "%s: residual epilogue needs residual/ldr", who);
    if (epilogue == WVD_EPI_BIAS_GATE_RES) WVD_REQUIRE(gate, "%s: gate epilogue needs gate", who);
    // developer experiments: bits 8..15 of `variant` override the rasterisation band width (0 = heuristic)
    const int band_override = (variant >> 8) & 0xff;
    variant &= 0xff;
    WVD_REQUIRE(variant >= WVD_GEMM_AUTO && variant <= WVD_GEMM_2CTA_M512, "%s: bad kernel variant %d", who, variant);

    Params p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.groups = groups;
    p.tiles_per_group = (int)((N + BN - 1) / BN);
    p.n_tiles = p.tiles_per_group * groups;
    p.k_blocks = (int)((K + BK - 1) / BK);
    for (int g = 0; g < MAX_GROUPS; ++g) p.bias[g] = (bias != nullptr && g < groups) ? (const __nv_bfloat16*)bias[g] : nullptr;
    p.gate = (const __nv_bfloat16*)gate;
    p.res = (const __nv_bfloat16*)residual; p.ldr = ldr;
    if (variant == WVD_GEMM_AUTO) variant = pick_variant(M, p.n_tiles, K);
    const int cg = variant == WVD_GEMM_1CTA ? 1 : 2;
    const int mt = variant == WVD_GEMM_2CTA_M512 ? 2 : 1;
    // Band width: N-tiles whose W rows are swept together while A streams past.  Measured (tools/gemm_sustained.py
    // --bands): 9-12 tiles is best at every c3 shape, also where the band (12 x 256 x 13,824 x 2 B = 85 MB) exceeds what
    // one would expect to stay L2-resident; narrower bands re-read A more often.  Bands are balanced over N.
    {
        const int nb = (p.n_tiles + 11) / 12;
        p.band = (p.n_tiles + nb - 1) / nb;
        if (band_override) p.band = band_override < p.n_tiles ? band_override : p.n_tiles;
    }
    const int unit_rows = BM * cg * mt;
    p.m_units = (int)((M + unit_rows - 1) / unit_rows);
    p.num_units = p.m_units * p.n_tiles;

    CUtensorMap tmA, tmB[MAX_GROUPS], tmC;
    int rc = get_tensor_map_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM * mt);
    if (rc) return rc;
    for (int g = 0; g < MAX_GROUPS; ++g) {
        rc = get_tensor_map_bf16(&tmB[g], W[g < groups ? g : 0], (uint64_t)N, (uint64_t)K, (uint64_t)ldw, BN / cg);
        if (rc) return rc;
    }
    rc = get_tensor_map_bf16(&tmC, Cp, (uint64_t)M, (uint64_t)(N * groups), (uint64_t)ldc, 32, CHUNK);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
#define WVD_GEMM_CASE(E)                                                                       \
    case E:                                                                                    \
        if (mt == 2) return launch<E, 2, 2>(tmA, tmB, tmC, p, s);                              \
        return cg == 2 ? launch<E, 2, 1>(tmA, tmB, tmC, p, s) : launch<E, 1, 1>(tmA, tmB, tmC, p, s);
    switch (epilogue) {
        WVD_GEMM_CASE(WVD_EPI_BIAS)
        WVD_GEMM_CASE(WVD_EPI_BIAS_GELU)
        WVD_GEMM_CASE(WVD_EPI_BIAS_RES)
        WVD_GEMM_CASE(WVD_EPI_BIAS_GATE_RES)
        WVD_GEMM_CASE(WVD_EPI_BIAS_MUL)
        WVD_GEMM_CASE(WVD_EPI_BIAS_GELU_T5)
        default: return set_error(WVD_ERR_INVALID, "%s: bad epilogue %d", who, epilogue);
    }
#undef WVD_GEMM_CASE
}

}  // namespace gemm

int gemm_read_diag(unsigned long long* out) {
    if (cudaMemcpyFromSymbol(out, g_diag, sizeof(unsigned long long) * 8) != cudaSuccess) return -1;
    unsigned long long zero[8] = {0};
    cudaMemcpyToSymbol(g_diag, zero, sizeof(zero));
    return 0;
}

}  // namespace wvd

extern "C" __attribute__((visibility("default"))) int wvd_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* C,
                             int64_t ldc, int64_t M, int64_t N, int64_t K, int epilogue, const void* gate,
                             const void* residual, int64_t ldr, wvd_stream_t stream) {
    const void* w1[1] = {W};
    const void* b1[1] = {bias};
    return wvd::gemm::run(A, lda, w1, ldw, bias ? b1 : nullptr, C, ldc, M, N, K, 1, epilogue, gate, residual, ldr,
                          WVD_GEMM_AUTO, stream, "wvd_gemm_bf16");
}

extern "C" __attribute__((visibility("default"))) int wvd_gemm_bf16_select(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* C,
                                    int64_t ldc, int64_t M, int64_t N, int64_t K, int epilogue, const void* gate,
                                    const void* residual, int64_t ldr, int variant, wvd_stream_t stream) {
    const void* w1[1] = {W};
    const void* b1[1] = {bias};
    return wvd::gemm::run(A, lda, w1, ldw, bias ? b1 : nullptr, C, ldc, M, N, K, 1, epilogue, gate, residual, ldr,
                          variant, stream, "wvd_gemm_bf16_select");
}

extern "C" __attribute__((visibility("default"))) int wvd_gemm_bf16_grouped(const void* A, int64_t lda, const void* const* W, int64_t ldw,
                                     const void* const* bias, void* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                                     int groups, int variant, wvd_stream_t stream) {
    return wvd::gemm::run(A, lda, W, ldw, bias, C, ldc, M, N, K, groups, WVD_EPI_BIAS, nullptr, nullptr, 0, variant, stream,
                          "wvd_gemm_bf16_grouped");
}
