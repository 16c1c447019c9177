"""Tensor-level wrappers over the C ABI (include/wvd.h).  PyTorch supplies device memory and the stream only.

All activations are 2-D row-major views (tokens, channels) whose last stride is 1; the row stride is passed as
the leading dimension, so column slices of fused buffers (q|k|v) are used without copies.
bf16 tensors run the tcgen05 kernels; fp32 tensors run the fp32 CUDA-core kernels (parity mode).  Anything else,
or a CPU tensor, raises: the hot path has no fallback.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import EPI_BIAS, EPI_BIAS_GATE_RES, EPI_BIAS_GELU, EPI_BIAS_GELU_T5, EPI_BIAS_MUL, EPI_BIAS_RES, WVD_BF16, WVD_F32, WvdError, check

Tensor = torch.Tensor

# instrumentation read by bench.py: number of C-ABI kernel launches, and (optionally) CUDA-event pairs around the
# dominant kernel (self-attention) recorded on the launching stream.
LAUNCHES = 0
PROFILE = None      # set to {"self_attention": []} to record (start, end) events


def _dt(t: Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return WVD_BF16
    if t.dtype == torch.float32:
        return WVD_F32
    raise WvdError(f"unsupported dtype {t.dtype}: the wvd kernels take bfloat16 (tensor-core path) or float32 (parity mode)")


def _chk2d(t: Tensor, name: str) -> Tensor:
    if not t.is_cuda:
        raise WvdError(f"{name} is on {t.device}: the wvd hot path runs on CUDA (sm_100a) only; there is no CPU fallback")
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise WvdError(f"{name} must be a 2-D view with unit last stride, got shape {tuple(t.shape)} strides {t.stride()}")
    return t


def _ld(t: Tensor) -> int:
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def _vec(t: Optional[Tensor], n: int, name: str, like: Tensor):
    if t is None:
        return None
    t = t.reshape(-1)
    if t.numel() != n or t.dtype != like.dtype or not t.is_cuda or t.stride(0) != 1:
        raise WvdError(f"{name} must be a contiguous {like.dtype} CUDA vector of length {n}, got {tuple(t.shape)} {t.dtype}")
    return t


def _p(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    global LAUNCHES
    LAUNCHES += 1
    return torch.cuda.current_stream().cuda_stream


def as_2d(x: Tensor) -> Tensor:
    """(1, N, D) or (N, D) -> (N, D) view."""
    if x.dim() == 3:
        if x.shape[0] != 1:
            raise WvdError("batched activations must be looped over by the caller (batch 1 per call)")
        return x[0]
    return x


def ln_modulate(x: Tensor, shift: Optional[Tensor] = None, scale: Optional[Tensor] = None,
                weight: Optional[Tensor] = None, bias: Optional[Tensor] = None, eps: float = 1e-6,
                out: Optional[Tensor] = None) -> Tensor:
    """LN(x)*(1+scale)+shift, or LN(x)*weight+bias (wan_video_dit.py:64-65, 206-208)."""
    x = _chk2d(x, "x")
    n, d = x.shape
    if out is None:
        out = torch.empty((n, d), dtype=x.dtype, device=x.device)
    _chk2d(out, "out")
    shift, scale = _vec(shift, d, "shift", x), _vec(scale, d, "scale", x)
    weight, bias = _vec(weight, d, "weight", x), _vec(bias, d, "bias", x)
    if n == 0:
        return out
    check(_lib.load().wvd_ln_modulate(x.data_ptr(), _ld(x), _p(shift), _p(scale), _p(weight), _p(bias), out.data_ptr(),
                                      _ld(out), n, d, eps, _dt(x), _stream()), "wvd_ln_modulate")
    return out


def make_rope_table(freqs: Tuple[Tensor, Tensor, Tensor], device) -> Tensor:
    """The model's three complex128 tables (wan_video_dit.py:75-89) -> one fp32 (3, 1024, 32, 2) cos/sin table."""
    tab = torch.zeros(3, 1024, 32, 2, dtype=torch.float32)
    for a, f in enumerate(freqs):
        n, w = f.shape
        if n > 1024 or w > 32:
            raise WvdError(f"rope table axis {a} has shape {tuple(f.shape)}; expected <= (1024, 32)")
        tab[a, :n, :w, 0] = f.real.to(torch.float32)
        tab[a, :n, :w, 1] = f.imag.to(torch.float32)
    return tab.to(device).contiguous()


def rope_table_from_freqs(freqs: Tensor, device) -> Tensor:
    """The reference's per-token ``freqs`` tensor -- (N, 1, 64) complex, what DiTBlock.forward / rope_apply take
    (wan_video_dit.py:92-97, wan_video_new.py:1392-1396) -- as the (N, 64, 2) fp32 (cos, sin) table of the kernel's
    per-token mode."""
    if not torch.is_complex(freqs) or freqs.shape[-1] != 64:
        raise WvdError(f"freqs must be a complex (N, 1, 64) tensor, got {tuple(freqs.shape)} {freqs.dtype}")
    f = freqs.reshape(-1, 64)
    return torch.stack([f.real, f.imag], dim=-1).to(device=device, dtype=torch.float32).contiguous()


def qk_rmsnorm_rope(q: Tensor, k: Optional[Tensor], wq: Tensor, wk: Optional[Tensor], eps: float,
                    rope_table: Optional[Tensor] = None, grid: Tuple[int, int, int] = (1, 1, 1),
                    token_offset: int = 0, frame_ids: Optional[Tensor] = None,
                    q_out: Optional[Tensor] = None, k_out: Optional[Tensor] = None) -> Tuple[Tensor, Optional[Tensor]]:
    """q <- rope(rmsnorm(q)*wq), k likewise; in place unless q_out/k_out given (wan_video_dit.py:92-111,141-145)."""
    q = _chk2d(q, "q")
    n, d = q.shape
    q_out = q if q_out is None else _chk2d(q_out, "q_out")
    wq = _vec(wq, d, "norm_q.weight", q)
    if k is not None:
        k = _chk2d(k, "k")
        if k.shape != q.shape or k.dtype != q.dtype:
            raise WvdError("q and k must have the same shape/dtype (process the cross-attention k separately)")
        k_out = k if k_out is None else _chk2d(k_out, "k_out")
        wk = _vec(wk, d, "norm_k.weight", q)
    if rope_table is not None and tuple(grid) == (0, 0, 0):
        if rope_table.dtype != torch.float32 or tuple(rope_table.shape) != (n, 64, 2) or not rope_table.is_cuda or not rope_table.is_contiguous():
            raise WvdError("per-token rope table must be a contiguous (n_tokens, 64, 2) fp32 CUDA tensor (see rope_table_from_freqs)")
        if frame_ids is not None or token_offset != 0:
            raise WvdError("per-token rope table takes no frame_ids / token_offset")
    elif rope_table is not None:
        if rope_table.dtype != torch.float32 or tuple(rope_table.shape) != (3, 1024, 32, 2) or not rope_table.is_cuda:
            raise WvdError("rope_table must be the (3,1024,32,2) fp32 CUDA table from make_rope_table()")
        if frame_ids is not None and (frame_ids.dtype != torch.int32 or not frame_ids.is_cuda):
            raise WvdError("frame_ids must be an int32 CUDA tensor")
    gf, gh, gw = grid
    if n == 0:
        return q_out, k_out
    check(_lib.load().wvd_qk_rmsnorm_rope(
        q.data_ptr(), _ld(q), _p(k), _ld(k) if k is not None else 0, wq.data_ptr(), _p(wk), q_out.data_ptr(), _ld(q_out),
        _p(k_out) if k is not None else None, _ld(k_out) if k is not None else 0, n, d, 128, eps, _p(rope_table),
        _p(frame_ids), gf, gh, gw, token_offset, _dt(q), _stream()), "wvd_qk_rmsnorm_rope")
    return q_out, k_out


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, epilogue: int = EPI_BIAS,
           gate: Optional[Tensor] = None, residual: Optional[Tensor] = None, out: Optional[Tensor] = None,
           variant: int = _lib.GEMM_AUTO) -> Tensor:
    """out = epilogue(x @ weight.T + bias) -- F.linear with fused GELU-tanh / residual / gate*y + residual.
    ``variant`` names the 1-CTA or the CTA-pair kernel explicitly (tests, tools); the default picks by tile-wave model."""
    x = _chk2d(x, "x")
    weight = _chk2d(weight, "weight")
    m, k = x.shape
    n, k2 = weight.shape
    if k2 != k or weight.dtype != x.dtype:
        raise WvdError(f"linear: x {tuple(x.shape)} {x.dtype} vs weight {tuple(weight.shape)} {weight.dtype}")
    if out is None:
        out = torch.empty((m, n), dtype=x.dtype, device=x.device)
    _chk2d(out, "out")
    if out.shape != (m, n):
        raise WvdError(f"linear: out has shape {tuple(out.shape)}, expected {(m, n)}")
    bias = _vec(bias, n, "bias", x)
    gate = _vec(gate, n, "gate", x)
    if residual is not None:
        residual = _chk2d(residual, "residual")
        if residual.shape != (m, n) or residual.dtype != x.dtype:
            raise WvdError("linear: residual shape/dtype mismatch")
    _dt(x)
    if m == 0:
        return out
    args = (x.data_ptr(), _ld(x), weight.data_ptr(), _ld(weight), _p(bias), out.data_ptr(), _ld(out), m, n, k, epilogue,
            _p(gate), _p(residual), _ld(residual) if residual is not None else 0)
    if x.dtype != torch.bfloat16:
        check(_lib.load().wvd_gemm_f32(*args, _stream()), "wvd_gemm_f32")
    elif variant == _lib.GEMM_AUTO:
        check(_lib.load().wvd_gemm_bf16(*args, _stream()), "wvd_gemm_bf16")
    else:
        check(_lib.load().wvd_gemm_bf16_select(*args, variant, _stream()), "wvd_gemm_bf16_select")
    return out


def linear_grouped(x: Tensor, weights, biases, out: Tensor, variant: int = _lib.GEMM_AUTO) -> Tensor:
    """out[:, g*N:(g+1)*N] = x @ weights[g].T + biases[g] for up to 3 same-shape weights in ONE launch (the q | k | v
    projections, wan_video_dit.py:131-133).  fp32 parity mode, or shapes the grouped kernel does not take, run the
    projections one by one through ``linear`` (same kernels, same results)."""
    x = _chk2d(x, "x")
    out = _chk2d(out, "out")
    g = len(weights)
    m, k = x.shape
    n = weights[0].shape[0]
    if out.shape != (m, g * n):
        raise WvdError(f"linear_grouped: out has shape {tuple(out.shape)}, expected {(m, g * n)}")
    same = all(_chk2d(w, "weight").shape == (n, k) and w.dtype == x.dtype and _ld(w) == _ld(weights[0]) for w in weights)
    if x.dtype != torch.bfloat16 or not same or g > 3 or n % 64 != 0:
        for i, (w, b) in enumerate(zip(weights, biases)):
            linear(x, w, b, out=out[:, i * n:(i + 1) * n], variant=variant)
        return out
    biases = [_vec(b, n, "bias", x) for b in biases]
    if m == 0:
        return out
    import ctypes
    wp = (ctypes.c_void_p * 3)(*[w.data_ptr() for w in weights])
    bp = (ctypes.c_void_p * 3)(*[(b.data_ptr() if b is not None else None) for b in biases])
    check(_lib.load().wvd_gemm_bf16_grouped(x.data_ptr(), _ld(x), wp, _ld(weights[0]), bp, out.data_ptr(), _ld(out), m, n, k,
                                            g, variant, _stream()), "wvd_gemm_bf16_grouped")
    return out


def attention(q: Tensor, k: Tensor, v: Tensor, num_heads: int, out: Optional[Tensor] = None,
              scale: Optional[float] = None, kernel: int = _lib.ATTN_AUTO) -> Tensor:
    """softmax(q k^T / sqrt(128)) v per head; q (Sq, H*128), k/v (Sk, H*128) views (wan_video_dit.py:28-61).
    ``kernel`` names one of the two bf16 kernels explicitly (parity tests); the default picks by key length."""
    q, k, v = _chk2d(q, "q"), _chk2d(k, "k"), _chk2d(v, "v")
    sq, width = q.shape
    sk = k.shape[0]
    if width != num_heads * 128 or k.shape[1] != width or v.shape != k.shape or k.dtype != q.dtype or v.dtype != q.dtype:
        raise WvdError(f"attention: head_dim must be 128 and q/k/v consistent (q {tuple(q.shape)}, k {tuple(k.shape)}, "
                       f"v {tuple(v.shape)}, heads {num_heads})")
    if out is None:
        out = torch.empty((sq, width), dtype=q.dtype, device=q.device)
    _chk2d(out, "out")
    scale = 1.0 / math.sqrt(128.0) if scale is None else scale
    _dt(q)
    if sq == 0:
        return out
    if sk == 0:
        raise WvdError("attention: empty key/value sequence")
    prof = PROFILE is not None and sq == sk
    if prof:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    args = (q.data_ptr(), _ld(q), k.data_ptr(), _ld(k), v.data_ptr(), _ld(v), out.data_ptr(), _ld(out), num_heads, sq,
            sk, 128, scale)
    if q.dtype != torch.bfloat16:
        check(_lib.load().wvd_attention_fwd_f32(*args, _stream()), "wvd_attention_fwd_f32")
    elif kernel == _lib.ATTN_AUTO:
        check(_lib.load().wvd_attention_fwd(*args, _stream()), "wvd_attention_fwd")
    else:
        check(_lib.load().wvd_attention_fwd_select(*args, kernel, _stream()), "wvd_attention_fwd_select")
    if prof:
        e1.record()
        PROFILE.setdefault("self_attention", []).append((e0, e1, num_heads, sq))
    return out


def scale_add(x: Tensor, y: Tensor, scale: float, out: Optional[Tensor] = None) -> Tensor:
    """x + y*scale (VACE hint injection, wan_video_new.py:1445-1450).  Contiguous tensors of equal shape."""
    if not (x.is_cuda and y.is_cuda) or x.shape != y.shape or x.dtype != y.dtype or not x.is_contiguous() or not y.is_contiguous():
        raise WvdError("scale_add: x and y must be contiguous CUDA tensors of equal shape/dtype")
    out = torch.empty_like(x) if out is None else out
    if x.numel() == 0:
        return out
    check(_lib.load().wvd_scale_add(x.data_ptr(), y.data_ptr(), float(scale), out.data_ptr(), x.numel(), _dt(x), _stream()),
          "wvd_scale_add")
    return out


def cfg_euler_step(x: Tensor, v_posi: Tensor, v_nega: Optional[Tensor], cfg_scale: float, dsigma: float,
                   out: Optional[Tensor] = None) -> Tensor:
    """x + (v_nega + cfg_scale * (v_posi - v_nega)) * dsigma in one pass, with the reference's per-operation rounding
    (wan_video_new.py:535,540; flow_match.py:72-82).  v_nega None: x + v_posi * dsigma."""
    ts = [t for t in (x, v_posi, v_nega) if t is not None]
    if any((not t.is_cuda) or t.shape != x.shape or t.dtype != x.dtype or not t.is_contiguous() for t in ts):
        raise WvdError("cfg_euler_step: contiguous CUDA tensors of equal shape/dtype expected")
    out = torch.empty_like(x) if out is None else out
    if x.numel() == 0:
        return out
    check(_lib.load().wvd_cfg_euler_step(x.data_ptr(), v_posi.data_ptr(), _p(v_nega), float(cfg_scale), float(dsigma),
                                         out.data_ptr(), x.numel(), _dt(x), _stream()), "wvd_cfg_euler_step")
    return out


def gate_residual(x: Tensor, gate: Tensor, y: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """x + gate*y (GateModule, wan_video_dit.py:189-194)."""
    x, y = _chk2d(x, "x"), _chk2d(y, "y")
    if not x.is_contiguous() or not y.is_contiguous() or x.shape != y.shape:
        raise WvdError("gate_residual: contiguous x/y of equal shape required")
    n, d = x.shape
    gate = _vec(gate, d, "gate", x)
    out = torch.empty_like(x) if out is None else out
    check(_lib.load().wvd_gate_residual(x.data_ptr(), gate.data_ptr(), y.data_ptr(), out.data_ptr(), n, d, _dt(x), _stream()),
          "wvd_gate_residual")
    return out


def ulysses_pack_qkv(qkv: Tensor, heads: int, world: int, out: Optional[Tensor] = None) -> Tensor:
    """(n_local, 3*heads*128) [q|k|v] -> (world, n_local, 3, heads/world, 128): dest-rank-major all-to-all send buffer."""
    qkv = _chk2d(qkv, "qkv")
    n = qkv.shape[0]
    if qkv.dtype != torch.bfloat16 or qkv.shape[1] != 3 * heads * 128:
        raise WvdError("ulysses_pack_qkv: bf16 (n, 3*heads*128) expected")
    if out is None:
        out = torch.empty((world, n, 3, heads // world, 128), dtype=qkv.dtype, device=qkv.device)
    check(_lib.load().wvd_ulysses_pack_qkv(qkv.data_ptr(), _ld(qkv), out.data_ptr(), n, heads, 128, world, _stream()),
          "wvd_ulysses_pack_qkv")
    return out


def ulysses_unpack_out(recv: Tensor, heads: int, world: int, out: Optional[Tensor] = None) -> Tensor:
    """(world, n_local, (heads/world)*128) -> (n_local, heads*128)."""
    if recv.dtype != torch.bfloat16 or not recv.is_contiguous() or not recv.is_cuda:
        raise WvdError("ulysses_unpack_out: contiguous bf16 CUDA tensor expected")
    n = recv.shape[1]
    if out is None:
        out = torch.empty((n, heads * 128), dtype=recv.dtype, device=recv.device)
    _chk2d(out, "out")
    check(_lib.load().wvd_ulysses_unpack_out(recv.data_ptr(), out.data_ptr(), _ld(out), n, heads, 128, world, _stream()),
          "wvd_ulysses_unpack_out")
    return out


def _ptr_array(ptrs):
    import ctypes
    arr = (ctypes.c_void_p * _lib.MAX_PEERS)()
    for i, v in enumerate(ptrs):
        arr[i] = int(v)
    return arr


def ulysses_scatter_qkv(qkv: Tensor, heads: int, recv_ptrs, rank: int) -> None:
    """Fused pack + all-to-all over NVLink: rank `rank`'s (n_local, 3*heads*128) q|k|v is stored straight into every
    rank's receive buffer (recv_ptrs: data pointers of a symmetric-memory rendezvous), in the layout its attention
    reads in place.  The caller issues a cross-rank barrier afterwards."""
    qkv = _chk2d(qkv, "qkv")
    world = len(recv_ptrs)
    if qkv.dtype != torch.bfloat16 or qkv.shape[1] != 3 * heads * 128 or world > _lib.MAX_PEERS:
        raise WvdError("ulysses_scatter_qkv: bf16 (n, 3*heads*128) and at most 8 ranks expected")
    check(_lib.load().wvd_ulysses_scatter_qkv(qkv.data_ptr(), _ld(qkv), _ptr_array(recv_ptrs), qkv.shape[0], heads, 128,
                                              world, rank, _stream()), "wvd_ulysses_scatter_qkv")


def ulysses_scatter_v(qkv: Tensor, heads: int, recv_ptrs, rank: int) -> None:
    """The v third of ``ulysses_scatter_qkv`` (q and k were stored into the peers by ``qk_rmsnorm_rope_scatter``)."""
    qkv = _chk2d(qkv, "qkv")
    world = len(recv_ptrs)
    if qkv.dtype != torch.bfloat16 or qkv.shape[1] != 3 * heads * 128 or world > _lib.MAX_PEERS:
        raise WvdError("ulysses_scatter_v: bf16 (n, 3*heads*128) and at most 8 ranks expected")
    check(_lib.load().wvd_ulysses_scatter_v(qkv.data_ptr(), _ld(qkv), _ptr_array(recv_ptrs), qkv.shape[0], heads, 128,
                                            world, rank, _stream()), "wvd_ulysses_scatter_v")


def qk_rmsnorm_rope_scatter(q: Tensor, k: Tensor, wq: Tensor, wk: Tensor, eps: float, rope_table: Tensor,
                            grid: Tuple[int, int, int], token_offset: int, frame_ids: Optional[Tensor], recv_ptrs,
                            rank: int) -> None:
    """``qk_rmsnorm_rope`` whose stores are the q / k part of the Ulysses exchange: every finished head goes straight
    into the owning rank's receive buffer (peer pointers of a symmetric-memory rendezvous); q and k are NOT modified."""
    q, k = _chk2d(q, "q"), _chk2d(k, "k")
    n, d = q.shape
    world = len(recv_ptrs)
    if q.dtype != torch.bfloat16 or k.shape != q.shape or k.dtype != q.dtype or world > _lib.MAX_PEERS:
        raise WvdError("qk_rmsnorm_rope_scatter: bf16 q / k of equal shape and at most 8 ranks expected")
    if rope_table.dtype != torch.float32 or tuple(rope_table.shape) != (3, 1024, 32, 2) or not rope_table.is_cuda:
        raise WvdError("rope_table must be the (3,1024,32,2) fp32 CUDA table from make_rope_table()")
    wq, wk = _vec(wq, d, "norm_q.weight", q), _vec(wk, d, "norm_k.weight", q)
    gf, gh, gw = grid
    check(_lib.load().wvd_qk_rmsnorm_rope_scatter(q.data_ptr(), _ld(q), k.data_ptr(), _ld(k), wq.data_ptr(), wk.data_ptr(),
                                                  _ptr_array(recv_ptrs), n, d, 128, eps, rope_table.data_ptr(), _p(frame_ids),
                                                  gf, gh, gw, token_offset, world, rank, _stream()), "wvd_qk_rmsnorm_rope_scatter")


def attention_scatter(q: Tensor, k: Tensor, v: Tensor, num_heads: int, out_ptrs, ldo: int, rows_per_peer: int,
                      col_offset: int, scale: Optional[float] = None, kernel: int = _lib.ATTN_AUTO) -> None:
    """ops.attention over the local heads whose epilogue stores query row t into rank t // rows_per_peer's
    (rows_per_peer, ldo) buffer at columns [col_offset, col_offset + num_heads*128) (peer pointers, NVLink): the
    Ulysses return all-to-all fused into the attention kernel.  The caller issues a cross-rank barrier afterwards."""
    q, k, v = _chk2d(q, "q"), _chk2d(k, "k"), _chk2d(v, "v")
    sq, width = q.shape
    sk = k.shape[0]
    if q.dtype != torch.bfloat16 or width != num_heads * 128 or k.shape[1] != width or v.shape != k.shape:
        raise WvdError("attention_scatter: bf16 q/k/v with head_dim 128 expected")
    if sq == 0:
        return
    scale = 1.0 / math.sqrt(128.0) if scale is None else scale
    prof = PROFILE is not None and sq == sk
    if prof:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(_lib.load().wvd_attention_fwd_scatter(q.data_ptr(), _ld(q), k.data_ptr(), _ld(k), v.data_ptr(), _ld(v),
                                                _ptr_array(out_ptrs), ldo, rows_per_peer, col_offset, len(out_ptrs),
                                                num_heads, sq, sk, 128, scale, kernel, _stream()), "wvd_attention_fwd_scatter")
    if prof:
        e1.record()
        PROFILE.setdefault("self_attention", []).append((e0, e1, num_heads, sq))


def attention_bias(q: Tensor, k: Tensor, v: Tensor, num_heads: int, bias: Optional[Tensor] = None,
                   key_mask: Optional[Tensor] = None, scale: float = 1.0, out: Optional[Tensor] = None) -> Tensor:
    """softmax(q k^T * scale + bias[h][j - i]; masked keys -> finfo.min) v for head_dim 64 -- the umT5 encoder's
    self-attention (wan_video_text_encoder.py:55-90).  q (Lq, H*64), k / v (Lk, H*64) views; bias (H, Lq + Lk - 1) in the
    activation dtype, indexed by (j - i) + Lq - 1; key_mask int32 (Lk,), 0 = masked."""
    q, k, v = _chk2d(q, "q"), _chk2d(k, "k"), _chk2d(v, "v")
    lq, width = q.shape
    lk = k.shape[0]
    if width != num_heads * 64 or k.shape[1] != width or v.shape != k.shape or k.dtype != q.dtype or v.dtype != q.dtype:
        raise WvdError(f"attention_bias: head_dim must be 64 and q/k/v consistent (q {tuple(q.shape)}, k {tuple(k.shape)}, heads {num_heads})")
    if bias is not None:
        bias = _chk2d(bias, "bias")
        if bias.shape[0] != num_heads or bias.shape[1] < lq + lk - 1 or bias.dtype != q.dtype:
            raise WvdError(f"attention_bias: bias must be ({num_heads}, >= {lq + lk - 1}) {q.dtype}, got {tuple(bias.shape)} {bias.dtype}")
    if key_mask is not None:
        key_mask = key_mask.reshape(-1)
        if key_mask.numel() != lk or key_mask.dtype != torch.int32 or not key_mask.is_cuda or key_mask.stride(0) != 1:
            raise WvdError("attention_bias: key_mask must be a contiguous int32 CUDA vector of length Lk")
    if out is None:
        out = torch.empty((lq, width), dtype=q.dtype, device=q.device)
    _chk2d(out, "out")
    if lq == 0:
        return out
    if lk == 0:
        raise WvdError("attention_bias: empty key/value sequence")
    check(_lib.load().wvd_attention_bias_fwd(q.data_ptr(), _ld(q), k.data_ptr(), _ld(k), v.data_ptr(), _ld(v), _p(bias),
                                             _ld(bias) if bias is not None else 0, _p(key_mask), out.data_ptr(), _ld(out),
                                             num_heads, lq, lk, 64, float(scale), _dt(q), _stream()), "wvd_attention_bias_fwd")
    return out


def editor_step(z_main: Tensor, z_edit: Tensor, v_posi, v_nega, frame_to_key: Tensor, key_idx: Tensor, cfg_scale: float,
                dt: float, alpha: float, beta: float, dsigma: float = 0.0, euler: bool = True):
    """The keyframe editor's per-step arithmetic in one kernel (wan_video_editor.py:107-165, 362-390).

    z_main (B, C, T, H, W), z_edit (B, C, K, H, W); v_posi / v_nega: either the joint (B, C, T+K, H, W) model output or a
    (v_main, v_edit) pair of separate tensors; v_nega None = no CFG.  Returns the new (z_main, z_edit) for ``euler``,
    else the corrected (v_main, v_edit)."""
    b, c, t, h, w = z_main.shape
    kf = z_edit.shape[2]
    if z_edit.shape != (b, c, kf, h, w) or z_edit.dtype != z_main.dtype:
        raise WvdError("editor_step: z_edit must be (B, C, K, H, W) of z_main's dtype")
    for name, x in (("z_main", z_main), ("z_edit", z_edit)):
        if not x.is_cuda or not x.is_contiguous():
            raise WvdError(f"editor_step: {name} must be a contiguous CUDA tensor (there is no CPU fallback)")
    hw = h * w

    def parts(v, name):
        if v is None:
            return None, None, 0, 0
        if isinstance(v, (tuple, list)):
            vm, ve = v
            if vm.shape != z_main.shape or ve.shape != z_edit.shape:
                raise WvdError(f"editor_step: {name} pair must match z_main / z_edit")
            tensors, strides = (vm, ve), (t * hw, kf * hw)
        else:
            if v.shape != (b, c, t + kf, h, w):
                raise WvdError(f"editor_step: {name} must be the joint (B, C, T+K, H, W) velocity, got {tuple(v.shape)}")
            tensors, strides = (v, v), ((t + kf) * hw, (t + kf) * hw)
        for x in tensors:
            if not x.is_cuda or not x.is_contiguous() or x.dtype != z_main.dtype:
                raise WvdError(f"editor_step: {name} must be contiguous CUDA tensors of z_main's dtype")
        if tensors[0] is tensors[1]:
            return v.data_ptr(), v.data_ptr() + t * hw * v.element_size(), strides[0], strides[1]
        return tensors[0].data_ptr(), tensors[1].data_ptr(), strides[0], strides[1]

    pm, pe, sm, se = parts(v_posi, "v_posi")
    nm, ne, sm2, se2 = parts(v_nega, "v_nega")
    if pm is None:
        raise WvdError("editor_step: v_posi is required")
    if nm is not None and (sm2, se2) != (sm, se):
        raise WvdError("editor_step: v_posi and v_nega must have the same layout")
    for name, x, n in (("frame_to_key", frame_to_key, t), ("key_idx", key_idx, kf)):
        if x.dtype != torch.int32 or not x.is_cuda or x.numel() != n or not x.is_contiguous():
            raise WvdError(f"editor_step: {name} must be a contiguous int32 CUDA vector of length {n}")
    out_main, out_edit = torch.empty_like(z_main), torch.empty_like(z_edit)
    check(_lib.load().wvd_editor_step(z_main.data_ptr(), z_edit.data_ptr(), pm, pe, nm, ne, sm, se, frame_to_key.data_ptr(),
                                      key_idx.data_ptr(), b * c, t, kf, hw, float(cfg_scale), float(dt), float(alpha), float(beta),
                                      float(dsigma), 1 if euler else 0, out_main.data_ptr(), out_edit.data_ptr(), _dt(z_main),
                                      _stream()), "wvd_editor_step")
    return out_main, out_edit


def tile_blend(values: Tensor, weight: Tensor, tile: Tensor, y0: int, x0: int, is_bound, border_width) -> None:
    """values[:, :, :, y0:y0+th, x0:x0+tw] += tile * mask ; weight[y0:.., x0:..] += mask  (wan_video_vae.py:1081-1152).
    values (1, C, T, H, W), tile (1, C, T, th, tw), weight (H, W); is_bound = (top, bottom, left, right) volume boundaries,
    border_width = (rows, columns) of the linear ramp on the other sides."""
    if values.dim() != 5 or tile.dim() != 5 or values.shape[0] != 1 or tile.shape[:3] != values.shape[:3]:
        raise WvdError(f"tile_blend: values (1, C, T, H, W) and tile (1, C, T, th, tw) expected, got {tuple(values.shape)} / {tuple(tile.shape)}")
    for name, x in (("values", values), ("weight", weight), ("tile", tile)):
        if not x.is_cuda or not x.is_contiguous() or x.dtype != values.dtype:
            raise WvdError(f"tile_blend: {name} must be a contiguous CUDA tensor of the values' dtype (there is no CPU fallback)")
    _, c, t, h, w = values.shape
    th, tw = tile.shape[3], tile.shape[4]
    if tuple(weight.shape) != (h, w):
        raise WvdError(f"tile_blend: weight must be the ({h}, {w}) plane, got {tuple(weight.shape)}")
    bounds = sum(1 << i for i, b in enumerate(is_bound) if b)
    check(_lib.load().wvd_tile_blend(values.data_ptr(), weight.data_ptr(), tile.data_ptr(), c * t, h, w, th, tw, int(y0), int(x0),
                                     bounds, int(border_width[0]), int(border_width[1]), _dt(values), _stream()), "wvd_tile_blend")


def tile_finalize(values: Tensor, weight: Tensor, clamp: Optional[Tuple[float, float]] = None) -> Tensor:
    """values / weight in place (weight broadcast over channels and frames), optionally clamped (wan_video_vae.py:1151-1153)."""
    if not values.is_cuda or not values.is_contiguous() or not weight.is_contiguous() or weight.dtype != values.dtype:
        raise WvdError("tile_finalize: contiguous CUDA tensors of one dtype expected")
    _, c, t, h, w = values.shape
    lo, hi = clamp if clamp is not None else (0.0, 0.0)
    check(_lib.load().wvd_tile_finalize(values.data_ptr(), weight.data_ptr(), c * t, h * w, 1 if clamp is not None else 0,
                                        float(lo), float(hi), _dt(values), _stream()), "wvd_tile_finalize")
    return values
