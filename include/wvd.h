/* wvd.h -- C ABI of libwvd.so: the B200-native (sm_100a) kernels behind the Wan2.1(-VACE) video-DiT denoising
 * forward of Ditto/Editto (wangshiwen-ai-hku/video-styler).
 *
 * Boundary rules:
 *   - plain extern "C"; no torch / ATen / pybind types.  Loaded with ctypes.
 *   - every function returns int: 0 = OK, < 0 = error code; text via wvd_last_error() (thread-local).
 *   - all buffers are caller-owned DEVICE pointers; no hidden allocation, no hidden synchronisation;
 *     every launch takes the caller's cudaStream_t (pass torch.cuda.current_stream().cuda_stream).
 *   - activations are row-major (tokens, channels) with an explicit leading dimension in ELEMENTS,
 *     so views into fused buffers (e.g. the q|k|v buffer) need no copy.
 *   - dtype enum: WVD_BF16 (storage bf16, math fp32) or WVD_F32.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference repo).
 */
#ifndef WVD_H_
#define WVD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* wvd_stream_t; /* cudaStream_t */

enum { WVD_OK = 0, WVD_ERR_INVALID = -1, WVD_ERR_CUDA = -2, WVD_ERR_UNSUPPORTED = -3 };
enum { WVD_BF16 = 0, WVD_F32 = 1 };

/* GEMM epilogues.  y = A.W^T + bias (fp32 accumulate, rounded to the storage dtype like F.linear).       */
enum {
    WVD_EPI_BIAS = 0,          /* C = y                                   nn.Linear (wan_video_dit.py:131-134) */
    WVD_EPI_BIAS_GELU = 1,     /* C = gelu_tanh(y)                        ffn.0 + nn.GELU('tanh') (:209-210)    */
    WVD_EPI_BIAS_RES = 2,      /* C = residual + y                        x + cross_attn(...) (:227); before_proj(c)+x (wan_video_vace.py:15) */
    WVD_EPI_BIAS_GATE_RES = 3, /* C = residual + gate[n] * y              GateModule (:189-194, :226, :229)     */
    WVD_EPI_BIAS_MUL = 4,      /* C = y * residual (elementwise)          T5 gated FFN fc1(x) * gelu(gate(x)) (wan_video_text_encoder.py:108) */
    WVD_EPI_BIAS_GELU_T5 = 5   /* C = 0.5 y (1 + tanh(c (y + 0.044715 y^3))) evaluated OP BY OP in the tensor dtype, the umT5 encoder's hand-written GELU (wan_video_text_encoder.py:16-20): seven roundings in bf16 */
};

/* ---- library info / diagnostics ------------------------------------------------------------------- */
const char* wvd_last_error(void);
int wvd_version(void);          /* 10000*major + 100*minor + patch */
int wvd_sm_arch(void);          /* 100: the only architecture this library is compiled for (sm_100a) */
/* Synchronises the device, copies out and clears the in-kernel watchdog record (out[0] = number of mbarrier waits
 * that timed out, out[1] = tag of the first one, out[2] = block, out[3] = thread).  A timed-out wait also TRAPS the
 * kernel, so the launch fails with a sticky CUDA error and this call (like every later one) returns WVD_ERR_CUDA.  */
int wvd_debug_flags(unsigned long long out[8]);
/* Build identity: "wvd <version> sm_100a src=<first 16 hex digits of the sha256 over csrc + include at build time>".
 * tests/test_abi.py compares it with the hash of the sources in the tree, so a stale prebuilt library is caught.   */
const char* wvd_build_info(void);

/* ---- K1/K2: LayerNorm (+ AdaLN modulate) ------------------------------------------------------------
 * out = LN(x) * (1 + scale) + shift            (weight == bias == NULL; shift/scale of length dim)
 * out = LN(x) * weight + bias                  (shift == scale == NULL)
 * Replaces nn.LayerNorm + modulate(): wan_video_dit.py:64-65, 206-208, 225, 227, 228; Head.forward :262-269.
 * Statistics in fp32, biased variance.  In bf16 mode the intermediate roundings of the reference's eager
 * bf16 expression (LN -> bf16, (1+scale) -> bf16, product -> bf16, sum -> bf16) are reproduced.           */
int wvd_ln_modulate(const void* x, int64_t ldx, const void* shift, const void* scale, const void* weight,
                    const void* bias, void* out, int64_t ldo, int64_t n_tokens, int dim, float eps, int dtype,
                    wvd_stream_t stream);

/* ---- K3/K4: full-width RMSNorm of q and k (+ 3-D RoPE) ------------------------------------------------
 * q <- rope(rmsnorm(q) * wq), k <- rope(rmsnorm(k) * wk); RMS over the whole hidden dim (all heads).
 * Replaces RMSNorm.forward + rope_apply: wan_video_dit.py:92-111, 141-145, 177-178 and the rank-sliced twin
 * diffsynth/distributed/xdit_context_parallel.py:27-40 (token_offset = rank * tokens_per_rank).
 * rope_cs == NULL disables RoPE (cross-attention).  k == NULL processes q only.
 * rope_cs: float2 (cos, sin) tables laid out [3][1024][32]: axis 0 = frame (22 pairs used), 1 = height (21),
 * 2 = width (21) -- the reference's complex128 tables (wan_video_dit.py:75-89) cast to fp32.
 * Token n (global index token_offset + local row) sits at (f, h, w) = (n / (gh*gw), (n / gw) % gh, n % gw);
 * frame_ids (int32[gf], may be NULL) replaces f by frame_ids[f] (rope_indices, wan_video_dit.py:378-384).
 * Rows whose global index lies past the grid are the zero padding of the last Ulysses shard
 * (wan_video_new.py:1414-1416): they are processed with the last frame's angles and never attended.
 * Per-token mode: grid_f == grid_h == grid_w == 0 makes rope_cs a (n_tokens, 64) float2 (cos, sin) table indexed by
 * the local row -- the `freqs` tensor (N, 1, 64) complex that the reference passes to DiTBlock.forward
 * (wan_video_dit.py:214-230), cast to fp32 pairs.
 * q_out/k_out may alias q/k.  head_dim must be 128.                                                      */
int wvd_qk_rmsnorm_rope(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* wq, const void* wk,
                        void* q_out, int64_t ldqo, void* k_out, int64_t ldko, int64_t n_tokens, int dim,
                        int head_dim, float eps, const void* rope_cs, const int32_t* frame_ids, int grid_f,
                        int grid_h, int grid_w, int64_t token_offset, int dtype, wvd_stream_t stream);

/* ---- elementwise residual forms that are not fused into a GEMM epilogue -------------------------------
 * out = x + y * scale   (VACE hint injection, diffsynth/pipelines/wan_video_new.py:1445-1450)            */
int wvd_scale_add(const void* x, const void* y, float scale, void* out, int64_t n_elems, int dtype,
                  wvd_stream_t stream);
/* One pass for the tail of a denoising step: CFG combine (diffsynth/pipelines/wan_video_new.py:535) and the Euler
 * update of FlowMatchScheduler.step (diffsynth/schedulers/flow_match.py:72-82):
 *   out = x + (v_nega + cfg_scale * (v_posi - v_nega)) * dsigma,  every operation rounded to the tensor dtype like the
 * reference's eager expressions.  v_nega == NULL (cfg_scale == 1): out = x + v_posi * dsigma.  out may alias x.     */
int wvd_cfg_euler_step(const void* x, const void* v_posi, const void* v_nega, float cfg_scale, float dsigma, void* out,
                       int64_t n_elems, int dtype, wvd_stream_t stream);
/* ---- VAE tile blending (WanVideoVAE.tiled_decode / tiled_encode, wan_video_vae.py:1081-1204) -----------------------
 * wvd_tile_blend: values[pl, y0:y0+th, x0:x0+tw] += tile[pl] * mask and weight[y0:.., x0:..] += mask for `planes` =
 * channels x frames planes of H x W (weight is ONE H x W plane: the mask does not depend on channel or frame).
 * mask = min(ramp_h, ramp_w) with the reference's linear ramps of `border_*` pixels on the sides that are not volume
 * boundaries; `bounds` bit 0 = top, 1 = bottom, 2 = left, 3 = right is a boundary.  wvd_tile_finalize: values /= weight
 * (clamped to [lo, hi] if clamp != 0).  Reference rounding per operation in `dtype`; everything stays in HBM (the
 * reference accumulates on the CPU).                                                                               */
int wvd_tile_blend(void* values, void* weight, const void* tile, int planes, int H, int W, int th, int tw, int y0, int x0,
                   int bounds, int border_h, int border_w, int dtype, wvd_stream_t stream);
int wvd_tile_finalize(void* values, const void* weight, int planes, int64_t hw, int clamp, float lo, float hi, int dtype,
                      wvd_stream_t stream);

/* ---- Keyframe editor step (wan_video_editor.py:107-165 compute_velocity_correction, :362-390 CFG + split + Euler) ----
 * One pass over the main latents (BC, T, HW) and the edited-keyframe latents (BC, K, HW): CFG combine (v_nega may be
 * NULL), velocity correction at the keyframe positions (alpha, beta, dt as in the reference), then either the Euler
 * update z + v * dsigma (euler = 1; flow_match.py:72-82) or the corrected velocities themselves (euler = 0), with the
 * reference's per-operation rounding in `dtype`.  Velocities are (BC, frames, HW) slabs with explicit BC strides in
 * elements (the joint (B, C, T+K, H, W) model output is read in place).  frame_to_key[t] = k or -1; key_idx[k] = t.   */
int wvd_editor_step(const void* z_main, const void* z_edit, const void* vp_main, const void* vp_edit, const void* vn_main,
                    const void* vn_edit, int64_t v_main_bc_stride, int64_t v_edit_bc_stride, const int* frame_to_key,
                    const int* key_idx, int bc, int t_frames, int k_frames, int64_t hw, float cfg_scale, float dt, float alpha,
                    float beta, float dsigma, int euler, void* out_main, void* out_edit, int dtype, wvd_stream_t stream);

/* ---- K11: umT5 self-attention, head_dim 64, additive relative-position bias + key mask (wan_video_text_encoder.py:55-90) --
 * out = softmax_j(q.k * scale + bias[h][(j - i) + lq - 1]; key_mask[j] == 0 -> finfo(dtype).min) v ; bias is a
 * (heads, >= lq + lk - 1) table in `dtype` with row stride ld_bias (NULL = no bias), key_mask int32[lk] or NULL; fp32
 * softmax; the reference's rounding points (scores, scores + bias, probabilities) are kept in bf16 mode.             */
int wvd_attention_bias_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                           const void* bias, int64_t ld_bias, const int* key_mask, void* out, int64_t ldo, int num_heads,
                           int64_t lq, int64_t lk, int head_dim, float scale, int dtype, wvd_stream_t stream);

/* out = x + gate[c] * y (GateModule, wan_video_dit.py:189-194) -- only used when the producer is not a GEMM */
int wvd_gate_residual(const void* x, const void* gate, const void* y, void* out, int64_t n_tokens, int dim,
                      int dtype, wvd_stream_t stream);

/* ---- K5-K7, K10: dense contraction on tcgen05 / TMEM, TMA-fed -----------------------------------------
 * C[M,N] = epilogue(A[M,K] . W[N,K]^T + bias[N]); A, W, C, residual bf16 row-major with leading dims
 * lda/ldw/ldc/ldr (elements, multiples of 8); bias/gate bf16 vectors of length N (bias may be NULL).
 * Replaces F.linear call sites wan_video_dit.py:131-134, 157-160, 209-210 and wan_video_vace.py:15,21.   */
int wvd_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* C,
                  int64_t ldc, int64_t M, int64_t N, int64_t K, int epilogue, const void* gate,
                  const void* residual, int64_t ldr, wvd_stream_t stream);

/* Three variants of the kernel sit behind it: WVD_GEMM_1CTA (128 x 256 tiles per CTA), WVD_GEMM_2CTA (256 x 256 tiles on
 * CTA pairs, tcgen05.mma cta_group::2, two accumulator stages) and WVD_GEMM_2CTA_M512 (512 x 256 tiles on CTA pairs, the
 * two halves filling all of TMEM: fewest operand bytes per flop).  WVD_GEMM_AUTO picks the one whose whole-wave unit
 * count over the 148 SMs models fastest, and is what wvd_gemm_bf16 uses.  The selector is an argument (parity tests and
 * the cuBLAS comparison run all of them on the same inputs); all variants accumulate in the same order and give
 * bit-identical results.                                                                                            */
enum { WVD_GEMM_AUTO = 0, WVD_GEMM_1CTA = 1, WVD_GEMM_2CTA = 2, WVD_GEMM_2CTA_M512 = 3 };
int wvd_gemm_bf16_select(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* C,
                         int64_t ldc, int64_t M, int64_t N, int64_t K, int epilogue, const void* gate,
                         const void* residual, int64_t ldr, int variant, wvd_stream_t stream);
/* Grouped launch: C[:, g*N:(g+1)*N] = A . W[g]^T + bias[g] for g < groups (<= 3) in ONE launch -- the q | k | v
 * projections of SelfAttention (wan_video_dit.py:131-133) share their input and write the column slices of one
 * (M, 3N) buffer; A is read once per tile from L2 and the three problems fill the machine's waves together.
 * W[g] are (N, K) row-major with the same ldw; bias may be NULL or hold NULL entries; N % 64 == 0 when groups > 1. */
int wvd_gemm_bf16_grouped(const void* A, int64_t lda, const void* const* W, int64_t ldw, const void* const* bias,
                          void* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int groups, int variant,
                          wvd_stream_t stream);

/* ---- K8/K9: flash-style attention forward, head_dim 128, non-causal, no mask ---------------------------
 * out[s, h*128:(h+1)*128] = softmax(q_h k_h^T * scale) v_h ; q/k/v/out are (tokens, heads*128) bf16 views with
 * leading dims in elements.  Replaces flash_attention(): wan_video_dit.py:28-61 (self: sq == sk; cross: sk = 512). */
int wvd_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                      void* out, int64_t ldo, int num_heads, int64_t sq, int64_t sk, int head_dim, float scale,
                      wvd_stream_t stream);
/* Four kernels sit behind it: WVD_ATTN_ONE_TILE (one 128-row Q tile per CTA, 256 TMEM columns, two CTAs per SM: the
 * 512-key text cross-attention, where prologue and epilogue are as long as the 4-step main loop), WVD_ATTN_TWO_TILE (two Q
 * tiles per CTA, one CTA per SM: medium key lengths), WVD_ATTN_CG2 (2-CTA clusters, one tcgen05.mma cta_group::2 stream
 * with M = 256 over the pair, K/V split over the pair, triple-buffered S in TMEM; long self-attention) and WVD_ATTN_PAIR
 * (its cta_group::1 predecessor: K/V multicast to both CTAs; kept for A/B).  WVD_ATTN_AUTO picks by key length (ONE_TILE
 * for sk <= 1024, CG2 for sk >= 2048) and is what wvd_attention_fwd uses.  The selector is an ARGUMENT -- the library
 * keeps no mutable dispatch state -- so that the parity tests can run all kernels on the same inputs.               */
enum { WVD_ATTN_AUTO = 0, WVD_ATTN_TWO_TILE = 1, WVD_ATTN_PAIR = 2, WVD_ATTN_CG2 = 3, WVD_ATTN_ONE_TILE = 4, WVD_ATTN_CG2_PERSISTENT = 5 };
/* Resident CTAs per SM of the short-key kernel `which` (WVD_ATTN_ONE_TILE: 2 by construction); < 0 = error. */
int wvd_debug_attention_resident_ctas(int which);
int wvd_attention_fwd_select(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                             void* out, int64_t ldo, int num_heads, int64_t sq, int64_t sk, int head_dim, float scale,
                             int which, wvd_stream_t stream);

/* ---- fp32 fall-through kernels (fp32 mode of the parity contract; CUDA-core math, not tuned) ----------- */
int wvd_gemm_f32(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* C, int64_t ldc,
                 int64_t M, int64_t N, int64_t K, int epilogue, const void* gate, const void* residual,
                 int64_t ldr, wvd_stream_t stream);
int wvd_attention_fwd_f32(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                          void* out, int64_t ldo, int num_heads, int64_t sq, int64_t sk, int head_dim, float scale,
                          wvd_stream_t stream);

/* ---- C1: Ulysses layout helpers (diffsynth/distributed/xdit_context_parallel.py:110-131) ---------------
 * pack:   qkv (n_local, 3*heads*128) [q|k|v]  ->  send (P, n_local, 3, heads/P, 128)  (dest-rank major)
 *         after all-to-all the receive buffer is (P*n_local, 3*(heads/P)*128): q|k|v views with ld = 3*(heads/P)*128.
 * The attention output (P*n_local, (heads/P)*128) is already the return-trip send layout; after the second
 * all-to-all the buffer (P, n_local, (heads/P)*128) is consumed by unpack -> (n_local, heads*128).          */
int wvd_ulysses_pack_qkv(const void* qkv, int64_t ld, void* send, int64_t n_local, int heads, int head_dim,
                         int world, wvd_stream_t stream);
int wvd_ulysses_unpack_out(const void* recv, void* out, int64_t ldo, int64_t n_local, int heads, int head_dim,
                           int world, wvd_stream_t stream);

/* ---- C2: Ulysses exchange fused into the producing kernels over NVLink peer memory ---------------------
 * Replaces the two all-to-alls of xfuser's SeqAllToAll4D (behind usp_attn_forward,
 * diffsynth/distributed/xdit_context_parallel.py:110-131) with direct peer stores: the pointers are those of a
 * symmetric-memory rendezvous (one buffer per rank, same size everywhere); the caller orders the steps with two
 * cross-rank barriers per attention (after the scatter, after the attention).
 *   scatter_qkv : the pack above, with each destination's chunk stored into THAT rank's receive buffer
 *                 recv_ptrs[dest][(rank*n_local + row)][3][heads/P][128]  (what its attention reads in place)
 *   attention_fwd_scatter : wvd_attention_fwd over the local heads and all tokens whose epilogue stores query row t
 *                 into out_ptrs[t / rows_per_peer] at local row t % rows_per_peer, columns
 *                 [col_offset, col_offset + num_heads*128) of a (rows_per_peer, ldo) buffer -- the o-projection input. */
#define WVD_MAX_PEERS 8
/*   qk_rmsnorm_rope_scatter : wvd_qk_rmsnorm_rope (bf16, RoPE on) whose STORES are the q / k part of scatter_qkv: the
 *                 finished vector of head h goes straight to recv_ptrs[h / (heads/P)], slot q (0) or k (1) of row
 *                 rank*n_local + row -- no pack / scatter pass over q and k ("K3 writing straight into the send
 *                 layout"); scatter_v then moves only the v third of the buffer.                                        */
int wvd_qk_rmsnorm_rope_scatter(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* wq, const void* wk,
                                void* const* recv_ptrs, int64_t n_local, int dim, int head_dim, float eps,
                                const void* rope_cs, const int32_t* frame_ids, int grid_f, int grid_h, int grid_w,
                                int64_t token_offset, int world, int rank, wvd_stream_t stream);
int wvd_ulysses_scatter_v(const void* qkv, int64_t ld, void* const* recv_ptrs, int64_t n_local, int heads, int head_dim,
                          int world, int rank, wvd_stream_t stream);
int wvd_ulysses_scatter_qkv(const void* qkv, int64_t ld, void* const* recv_ptrs, int64_t n_local, int heads,
                            int head_dim, int world, int rank, wvd_stream_t stream);
int wvd_attention_fwd_scatter(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                              void* const* out_ptrs, int64_t ldo, int64_t rows_per_peer, int64_t col_offset, int world,
                              int num_heads, int64_t sq, int64_t sk, int head_dim, float scale, int which,
                              wvd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* WVD_H_ */
