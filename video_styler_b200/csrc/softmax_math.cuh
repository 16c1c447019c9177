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

}  // namespace attn
}  // namespace wvd
