// Per-row softmax arithmetic of the attention kernel (one query row per thread), shared with the developer
// microbenchmarks under tools/microbench/.
#pragma once
#include "ptx.cuh"

namespace wvd {
namespace attn {

// exp2 of columns [C0, C1) of the NS score columns a thread holds of one S row, against the reference point folded into negm_2; returns the row sum of the
// chunk and leaves the bf16 pairs in pk[0 .. (C1-C0)/2).  The MUFU pipe takes one warp instruction per 8 cycles and
// ptxas fills the gaps with whatever independent work the basic block holds (scaling FMAs, the row maximum).
// EMU_OF_4: how many of every 4 consecutive column pairs take exp2 on the FMA pipes (polynomial) instead of MUFU.
template <int NS, int C0, int C1, int EMU_OF_4, bool WITH_MAX = false>
__device__ __forceinline__ float exp_chunk(const uint32_t (&s)[NS], uint32_t* pk, uint64_t sl2_2, uint64_t negm_2,
                                           float* mx = nullptr) {
    // WITH_MAX: mx[0..4) also accumulate max over ALL NS columns, NS/pairs columns per column pair of the chunk,
    // written out in the same loop so that the FMNMX instructions land in the issue slots between the exponentials.
    constexpr int NP = (C1 - C0) / 2;
    constexpr int MAX_COLS_PER_PAIR = NS / NP;
    static_assert(!WITH_MAX || (NS % NP == 0 && MAX_COLS_PER_PAIR % 2 == 0), "row max does not tile the chunk");
    uint64_t sum_a = f2_pack(0.f, 0.f), sum_b = f2_pack(0.f, 0.f);
#pragma unroll
    for (int c = C0; c < C1; c += 2) {
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
        if (WITH_MAX) {
            const int q = (c - C0) >> 1;
#pragma unroll
            for (int t = 0; t < MAX_COLS_PER_PAIR; t += 2) {
                const int mc = q * MAX_COLS_PER_PAIR + t;
                float& acc = mx[((mc >> 1) & 3)];
                acc = fmax3(acc, __uint_as_float(s[mc]), __uint_as_float(s[mc + 1]));
            }
        }
        if (c & 2) sum_b = f2_add(sum_b, f2_pack(p0, p1));
        else sum_a = f2_add(sum_a, f2_pack(p0, p1));
        pk[(c - C0) >> 1] = pack_bf16x2(p0, p1);
    }
    float a0, a1;
    f2_unpack(f2_add(sum_a, sum_b), a0, a1);
    return a0 + a1;
}

// max(seed, s[C0 .. C1)).  Every chain starts from the seed so that the whole computation depends on it: seeded with
// the running reference it cannot be hoisted out of the basic block that holds the MUFU stream.
template <int NS, int C0, int C1>
__device__ __forceinline__ float row_max(const uint32_t (&s)[NS], float seed) {
    float mx[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) mx[a] = fmax3(seed, __uint_as_float(s[C0 + 2 * a]), __uint_as_float(s[C0 + 2 * a + 1]));
#pragma unroll
    for (int c = C0 + 8; c < C1; c += 8) {
#pragma unroll
        for (int a = 0; a < 4; ++a) mx[a] = fmax3(mx[a], __uint_as_float(s[c + 2 * a]), __uint_as_float(s[c + 2 * a + 1]));
    }
    return fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
}

// registers -> TMEM, NCOL 32-bit columns (a multiple of 8)
template <int NCOL>
__device__ __forceinline__ void store_p(uint32_t taddr, const uint32_t* pk) {
#pragma unroll
    for (int c = 0; c + 32 <= NCOL; c += 32) tmem_st_32x32b_x32(taddr + c, pk + c);
    if ((NCOL % 32) / 16 != 0) tmem_st_32x32b_x16(taddr + (NCOL / 32) * 32, pk + (NCOL / 32) * 32);
    if (NCOL % 16 != 0) tmem_st_32x32b_x8(taddr + (NCOL / 16) * 16, pk + (NCOL / 16) * 16);
}


__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// Epilogue of the attention kernels: a warp holds 32 query rows (one per lane) x 64 output columns as fp32 in TMEM
// (o_tmem = lane quarter + first column); out = O * inv_l as bf16, 128 bytes per row.
// A direct store from the TMEM-load layout makes every warp store touch 32 different rows (32 wavefronts of 16 bytes per
// instruction).
// Instead the warp stages its 4 KB through shared memory (XOR-swizzled 16-byte chunks: conflict-free both ways) and
// writes 4 full 128-byte row segments per instruction (4 wavefronts): 8x fewer LSU cycles.
// row_ptr(rr) returns the global address of columns [0, 64) of the warp's row rr (0..31), or nullptr for rows past the end
// (evaluated by the lanes that STORE row rr, so the Ulysses per-peer address computation works row by row).
// stage: TMEM -> (x inv_l, bf16) -> the warp's 4 KB of shared memory; after it returns the warp's O columns are out of TMEM
__device__ __forceinline__ void stage_o_warp(uint32_t o_tmem, float inv_l, uint32_t stage, int lane) {
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(o_tmem + c * 32, o);
        tc_wait_ld();
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
            const int chunk = c * 4 + q4;
            st_shared_v4(stage + lane * 128 + ((chunk ^ (lane & 7)) << 4),
                         pack_bf16x2(__uint_as_float(o[q4 * 8 + 0]) * inv_l, __uint_as_float(o[q4 * 8 + 1]) * inv_l),
                         pack_bf16x2(__uint_as_float(o[q4 * 8 + 2]) * inv_l, __uint_as_float(o[q4 * 8 + 3]) * inv_l),
                         pack_bf16x2(__uint_as_float(o[q4 * 8 + 4]) * inv_l, __uint_as_float(o[q4 * 8 + 5]) * inv_l),
                         pack_bf16x2(__uint_as_float(o[q4 * 8 + 6]) * inv_l, __uint_as_float(o[q4 * 8 + 7]) * inv_l));
        }
    }
    __syncwarp();
}

// flush: the staged 32 rows x 128 bytes -> global, 4 whole row segments per warp instruction
template <typename RowPtr>
__device__ __forceinline__ void flush_o_warp(uint32_t stage, int lane, RowPtr row_ptr) {
    const int sub = lane >> 3, chunk = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + sub;
        const uint4 v = ld_shared_v4(stage + rr * 128 + ((chunk ^ (rr & 7)) << 4));
        __nv_bfloat16* dst = row_ptr(rr);
        if (dst != nullptr) *reinterpret_cast<uint4*>(dst + chunk * 8) = v;
    }
    __syncwarp();
}

template <typename RowPtr>
__device__ __forceinline__ void store_o_warp_coalesced(uint32_t o_tmem, float inv_l, uint32_t stage, int lane, RowPtr row_ptr) {
    stage_o_warp(o_tmem, inv_l, stage, lane);
    flush_o_warp(stage, lane, row_ptr);
}

}  // namespace attn
}  // namespace wvd
