// K5-K7/K10: C = epilogue(A . W^T + bias) on tcgen05 tensor cores (sm_100a).
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0 (1 thread)  TMA producer: A (128x64) and W (256x64) bf16 tiles, 128-byte swizzle, 4-stage mbarrier ring
//   warp 1 (1 thread)  MMA issuer : tcgen05.mma cta_group::1 kind::f16, 128x256x16, fp32 accumulators in TMEM
//   warp 2             TMEM allocator (512 columns = 2 accumulator stages of 256 columns)
//   warps 4-7          epilogue: tcgen05.ld (one accumulator row per thread) -> bias / GELU-tanh / gate / residual
//                      in fp32 with the reference's bf16 rounding points -> 16-byte global stores
// The epilogue of tile i overlaps the main loop of tile i+1 through the two TMEM accumulator stages.
// Tiles are rasterised in bands of 8 N-tiles so that the W band (<= 2048 x K) stays L2-resident while A streams.
// M/N/K tails rely on TMA out-of-bounds zero fill; stores are predicated.
//
// Replaces the F.linear call sites of the reference (diffsynth/models/wan_video_dit.py:131-134,157-160,209-210;
// wan_video_vace.py:15,21) and fuses GateModule (:189-194), the ungated residual (:227) and nn.GELU('tanh').
#include "host_utils.h"
#include "ptx.cuh"

namespace wvd {
namespace gemm {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_BYTES = BN * BK * 2;   // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NUM_THREADS = 256;
constexpr int BAND = 8;                // N-tiles per rasterisation band
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr uint32_t IDESC = make_idesc_bf16(BM, BN, 0, 0);

struct Params {
    int M, N, K;
    int m_tiles, n_tiles, num_tiles, k_blocks;
    __nv_bfloat16* C;
    long long ldc;
    const __nv_bfloat16* bias;
    const __nv_bfloat16* gate;
    const __nv_bfloat16* res;
    long long ldr;
};

__device__ __forceinline__ void tile_coords(const Params& p, int tile, int& m_blk, int& n_blk) {
    const int per_band = p.m_tiles * BAND;
    const int band = tile / per_band;
    const int within = tile - band * per_band;
    const int n_start = band * BAND;
    const int bw = min(BAND, p.n_tiles - n_start);
    m_blk = within / bw;
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
__device__ __forceinline__ void epilogue_store8(const Params& p, const uint32_t* acc, long long row, int col) {
    // acc: 8 fp32 accumulator values (as bits) for columns col..col+7 of `row`
    __nv_bfloat162 y[4];
    if (p.bias != nullptr) {
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(p.bias + col));
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
    if (EPI == WVD_EPI_BIAS_GATE_RES) {
        const uint4 g = __ldg(reinterpret_cast<const uint4*>(p.gate + col));
        const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) y[q] = __hmul2_rn(*reinterpret_cast<const __nv_bfloat162*>(&gw[q]), y[q]);
    }
    if (EPI == WVD_EPI_BIAS_RES || EPI == WVD_EPI_BIAS_GATE_RES) {
        const uint4 r = *reinterpret_cast<const uint4*>(p.res + row * p.ldr + col);
        const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) y[q] = __hadd2_rn(*reinterpret_cast<const __nv_bfloat162*>(&rw[q]), y[q]);
    }
    uint4 o;
    o.x = *reinterpret_cast<const uint32_t*>(&y[0]);
    o.y = *reinterpret_cast<const uint32_t*>(&y[1]);
    o.z = *reinterpret_cast<const uint32_t*>(&y[2]);
    o.w = *reinterpret_cast<const uint32_t*>(&y[3]);
    *reinterpret_cast<uint4*>(p.C + row * p.ldc + col) = o;
}

template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - raw_u32);
    const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + s * 8; };
    auto empty_bar = [&](int s) { return bar_base + 32 + s * 8; };
    auto tfull_bar = [&](int a) { return bar_base + 64 + a * 8; };
    auto tempty_bar = [&](int a) { return bar_base + 80 + a * 8; };
    const uint32_t tmem_slot = bar_base + 96;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * STAGE_BYTES + 96);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------ TMA producer ------------------------------
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                int m_blk, n_blk;
                tile_coords(p, tile, m_blk, n_blk);
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1, 0x100 + stage);
                    const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
                    mbar_expect_tx(full_bar(stage), STAGE_BYTES);
                    tma_load_2d(a_dst, &tmA, full_bar(stage), kb * BK, m_blk * BM);
                    tma_load_2d(a_dst + A_BYTES, &tmB, full_bar(stage), kb * BK, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------ MMA issuer ------------------------------
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1, 0x200 + acc);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(full_bar(stage), phase, 0x300 + stage);
                    tc_fence_after();
                    const uint32_t a_addr = smem_base + stage * STAGE_BYTES;
                    const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                        const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                        umma_ss(d_tmem, da, db, IDESC, (kb | k) != 0 ? 1u : 0u);
                    }
                    tc_commit(empty_bar(stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull_bar(acc));
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ------------------------------ epilogue ------------------------------
        const int q = warp - 4;   // == warp % 4: the TMEM lane quarter this warp may access
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            int m_blk, n_blk;
            tile_coords(p, tile, m_blk, n_blk);
            mbar_wait(tfull_bar(acc), acc_phase, 0x400 + acc);
            tc_fence_after();
            const long long row = static_cast<long long>(m_blk) * BM + q * 32 + lane;
            const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
            const bool row_ok = row < p.M;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(taddr + c * 32, v);
                tc_wait_ld();
                const int col0 = n_blk * BN + c * 32;
                if (row_ok) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int col = col0 + g * 8;
                        if (col < p.N) epilogue_store8<EPI>(p, v + g * 8, row, col);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int EPI>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const Params& p, cudaStream_t stream) {
    static unsigned long long configured = 0;
    if (first_use_on_current_device(&configured))
        WVD_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
    gemm_bf16_kernel<EPI><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmA, tmB, p);
    WVD_CHECK_CUDA(cudaGetLastError());
    return WVD_OK;
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
    using namespace wvd;
    WVD_REQUIRE(A && W && C, "wvd_gemm_bf16: null pointer");
    WVD_REQUIRE(M > 0 && N > 0 && K > 0, "wvd_gemm_bf16: empty problem M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    WVD_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "wvd_gemm_bf16: dimension too large");
    WVD_REQUIRE(K % 8 == 0 && N % 8 == 0, "wvd_gemm_bf16: K and N must be multiples of 8 (got K=%lld N=%lld)", (long long)K, (long long)N);
    WVD_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0 && lda >= K && ldw >= K && ldc >= N,
                "wvd_gemm_bf16: leading dims must be multiples of 8 and cover the row");
    WVD_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0) && ((uintptr_t)C % 16 == 0) &&
                ((uintptr_t)bias % 16 == 0) && ((uintptr_t)gate % 16 == 0) && ((uintptr_t)residual % 16 == 0),
                "wvd_gemm_bf16: pointers must be 16-byte aligned");
    WVD_REQUIRE(epilogue >= WVD_EPI_BIAS && epilogue <= WVD_EPI_BIAS_GATE_RES, "wvd_gemm_bf16: bad epilogue %d", epilogue);
    if (epilogue == WVD_EPI_BIAS_RES || epilogue == WVD_EPI_BIAS_GATE_RES)
        WVD_REQUIRE(residual && ldr % 8 == 0 && ldr >= N, "wvd_gemm_bf16: residual epilogue needs residual/ldr");
    if (epilogue == WVD_EPI_BIAS_GATE_RES) WVD_REQUIRE(gate, "wvd_gemm_bf16: gate epilogue needs gate");

    CUtensorMap tmA, tmB;
    int rc = get_tensor_map_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, gemm::BM);
    if (rc) return rc;
    rc = get_tensor_map_bf16(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, gemm::BN);
    if (rc) return rc;

    gemm::Params p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.m_tiles = (int)((M + gemm::BM - 1) / gemm::BM);
    p.n_tiles = (int)((N + gemm::BN - 1) / gemm::BN);
    p.num_tiles = p.m_tiles * p.n_tiles;
    p.k_blocks = (int)((K + gemm::BK - 1) / gemm::BK);
    p.C = (__nv_bfloat16*)C; p.ldc = ldc;
    p.bias = (const __nv_bfloat16*)bias;
    p.gate = (const __nv_bfloat16*)gate;
    p.res = (const __nv_bfloat16*)residual; p.ldr = ldr;
    cudaStream_t s = (cudaStream_t)stream;
    switch (epilogue) {
        case WVD_EPI_BIAS: return gemm::launch<WVD_EPI_BIAS>(tmA, tmB, p, s);
        case WVD_EPI_BIAS_GELU: return gemm::launch<WVD_EPI_BIAS_GELU>(tmA, tmB, p, s);
        case WVD_EPI_BIAS_RES: return gemm::launch<WVD_EPI_BIAS_RES>(tmA, tmB, p, s);
        default: return gemm::launch<WVD_EPI_BIAS_GATE_RES>(tmA, tmB, p, s);
    }
}
