"""Host-side orchestration of one DiT block / VACE block / head over the C-ABI kernels.

Everything here is duck-typed on the reference's module tree (attribute names ``self_attn.q``, ``norm_q``,
``modulation``, ``ffn[0]`` ...), so the same functions drive both this package's own modules
(``wan_video_dit.WanModel``) and a reference ``diffsynth`` pipeline patched by ``install()``.
Weights are fetched through ``_lin`` / ``_param`` which unwrap ``diffsynth.vram_management`` wrappers
(``AutoWrappedModule.module``); nothing is copied or re-packed, so LoRA merges and ``load_state_dict`` keep working.

Data layout: activations are (tokens, channels) row-major bf16 in HBM; the residual stream ``x`` is updated in
place by the GEMM epilogues (o-proj: x += gate*y, cross o-proj: x += y, ffn.2: x += gate*y).
q|k|v live in ONE (tokens, 3*D) buffer: three GEMMs write its column slices, the RMSNorm+RoPE kernel updates
q and k in place, and the attention kernel reads the slices through strided TMA maps.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch

from . import ops as _cuda_ops

Tensor = torch.Tensor


@dataclass
class RopeInfo:
    """3-D RoPE addressing for the fused QK-RMSNorm+RoPE kernel (wan_video_new.py:1392-1396)."""
    table: Tensor                      # (3, 1024, 32, 2) fp32 cos/sin; per-token mode: (N, 64, 2) with grid (0, 0, 0)
    grid: Tuple[int, int, int]         # (f, h, w) token grid
    token_offset: int = 0              # global index of local row 0 (Ulysses shard)
    frame_ids: Optional[Tensor] = None  # int32 (f,) -- rope_indices (wan_video_dit.py:378-384)

    @staticmethod
    def from_freqs(freqs: Tensor, device, ops=_cuda_ops) -> "RopeInfo":
        """The reference's per-token ``freqs`` argument -- (N, 1, 64) complex, assembled at wan_video_new.py:1392-1396
        and handed to ``block(x, context, t_mod, freqs)`` -- for callers that drive blocks directly."""
        return RopeInfo(ops.rope_table_from_freqs(freqs, device), (0, 0, 0))


@dataclass
class Workspace:
    """Caller-owned scratch for one (tokens, dim, ffn) problem; reused by every block of every step, so the TMA
    tensor-map cache inside libwvd.so always hits."""
    n: int
    dim: int
    ffn: int
    ctx_len: int
    dtype: torch.dtype
    device: torch.device
    bufs: Dict[str, Tensor] = field(default_factory=dict)

    def get(self, name: str, shape) -> Tensor:
        t = self.bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=self.dtype, device=self.device)
            self.bufs[name] = t
        return t


_workspaces: Dict[tuple, Workspace] = {}


def workspace(n: int, dim: int, ffn: int, ctx_len: int, dtype, device) -> Workspace:
    key = (n, dim, ffn, ctx_len, dtype, str(device))
    ws = _workspaces.get(key)
    if ws is None:
        if len(_workspaces) > 8:
            _workspaces.clear()
        ws = Workspace(n, dim, ffn, ctx_len, dtype, torch.device(device))
        _workspaces[key] = ws
    return ws


def _unwrap(m):
    """diffsynth.vram_management.AutoWrappedModule keeps the real module in ``.module`` (layers.py:36-60)."""
    inner = getattr(m, "module", None)
    return inner if isinstance(inner, torch.nn.Module) and not hasattr(m, "weight") else m


def _lin(m, dtype, device):
    m = _unwrap(m)
    w, b = m.weight, m.bias
    if w.dtype != dtype or w.device != device:          # offloaded / differently typed weights: cast like AutoWrappedLinear
        w = w.to(device=device, dtype=dtype)
        b = None if b is None else b.to(device=device, dtype=dtype)
    return w, b


def _param(p, dtype, device):
    if p.dtype != dtype or p.device != device:
        p = p.to(device=device, dtype=dtype)
    return p


def _norm_w(m, dtype, device):
    return _param(_unwrap(m).weight, dtype, device)


def patch_embed(conv, x: Tensor, ops=_cuda_ops) -> Tensor:
    """Conv3d with kernel == stride (WanModel.patchify, wan_video_dit.py:339-345; VaceWanModel, wan_video_vace.py:58-59)
    as im2col + the GEMM kernel: (1, C, F, H, W) -> (F*H'*W', D) tokens in (f, h, w) order.  No cuDNN/TF32 involved."""
    conv = _unwrap(conv)
    pt, ph, pw = conv.kernel_size
    if tuple(conv.stride) != (pt, ph, pw) or x.shape[0] != 1:
        raise NotImplementedError("patch embedding expects stride == kernel_size and batch 1")
    _, c, f, h, w = x.shape
    cols = (x.reshape(c, f // pt, pt, h // ph, ph, w // pw, pw).permute(1, 3, 5, 0, 2, 4, 6)
            .reshape((f // pt) * (h // ph) * (w // pw), c * pt * ph * pw))
    wgt = _param(conv.weight, x.dtype, x.device).reshape(conv.weight.shape[0], -1)
    return ops.linear(cols.contiguous(), wgt, _param(conv.bias, x.dtype, x.device))


def block_modulation(block, t_mod: Tensor) -> Tensor:
    """(modulation + t_mod) -> (6, D) rows [shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp]
    (wan_video_dit.py:218-219).  Per-token modulation (4-D t_mod, Wan2.2 TI2V) is not on this path."""
    if t_mod.dim() != 3 or t_mod.shape[0] != 1:
        raise NotImplementedError("per-token / batched t_mod (seperated_timestep) is not supported by the wvd path")
    return (block.modulation.to(dtype=t_mod.dtype, device=t_mod.device) + t_mod)[0].contiguous()


def as_rope_info(freqs, device, ops=_cuda_ops) -> RopeInfo:
    """``freqs`` as the blocks' callers pass it: an engine.RopeInfo (model_fn / WanModel.rope_info) or the reference's
    complex per-token tensor."""
    if isinstance(freqs, RopeInfo):
        return freqs
    if torch.is_tensor(freqs) and torch.is_complex(freqs):
        return RopeInfo.from_freqs(freqs, device, ops)
    raise TypeError("freqs must be an engine.RopeInfo (WanModel.rope_info) or the reference's complex (N, 1, 64) tensor")


class TextCache:
    """Text-side work that is constant across the denoising loop, computed once per prompt instead of once per call:
    ``dit.text_embedding(context)`` (wan_video_new.py:1357) and every block's cross-attention K = norm_k(k(ctx)),
    V = v(ctx) (wan_video_dit.py:177-179) -- the reference recomputes them 100x per video (50 steps x posi / nega).
    They depend only on ``context`` and the weights, so outputs are bit-identical with and without the cache.

    Keyed by the context tensor OBJECT (held weakly) and its in-place version counter; entries are validated against
    the data pointer / version of the weights they were computed from, so a LoRA merge or ``load_state_dict`` after
    the fact simply misses.  Memory: 2 x (L_ctx x D) per block and prompt (0.5 GB per prompt for the 14B + VACE model)."""

    def __init__(self, max_prompts: int = 4):
        self.max_prompts = max_prompts
        self._entries = {}          # id(context) -> dict(ref, version, emb, kv={id(block): (kc, vc, stamp)})
        self.hits = self.misses = 0

    def clear(self):
        self._entries.clear()

    def entry(self, context: Tensor):
        import weakref
        e = self._entries.get(id(context))
        if e is not None and (e["ref"]() is not context or e["version"] != context._version):
            e = None
        if e is None:
            if len(self._entries) >= self.max_prompts:
                self._entries.pop(next(iter(self._entries)))
            e = dict(ref=weakref.ref(context), version=context._version, emb=None, emb_stamp=None, kv={})
            self._entries[id(context)] = e
        return e

    @staticmethod
    def stamp(*params):
        return tuple((p.data_ptr(), p._version) for p in params)


class SelfAttnExchange:
    """Single-GPU: attention reads q|k|v in place.  ulysses.UlyssesExchange overrides this with the all-to-all."""
    world = 1

    def attend(self, ops, qkv: Tensor, heads: int, out: Tensor, ws: Workspace) -> Tensor:
        d = heads * 128
        return ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], heads, out=out)

    def norm_rope_attend(self, ops, qkv: Tensor, heads: int, wq: Tensor, wk: Tensor, eps: float, rope: "RopeInfo",
                         out: Tensor, ws: Workspace) -> Tensor:
        """Full-width QK-RMSNorm + RoPE (before any head scatter, wan_video_dit.py:141-145 /
        xdit_context_parallel.py:110-117), then attention.  The fused Ulysses exchange overrides this to let the
        RoPE kernel store q and k straight into the peers' receive buffers."""
        d = heads * 128
        ops.qk_rmsnorm_rope(qkv[:, :d], qkv[:, d:2 * d], wq, wk, eps, rope.table, rope.grid, rope.token_offset, rope.frame_ids)
        return self.attend(ops, qkv, heads, out, ws)


_LOCAL = SelfAttnExchange()


def dit_block_forward(block, x: Tensor, context: Tensor, t_mod: Tensor, rope: RopeInfo, ws: Workspace,
                      ops=_cuda_ops, exchange: SelfAttnExchange = _LOCAL, text_entry: Optional[dict] = None) -> Tensor:
    """DiTBlock.forward (wan_video_dit.py:214-230) on a (tokens, D) residual stream, updated IN PLACE.

    context: (L_ctx, D) text embedding (already through dit.text_embedding).  text_entry: a TextCache entry of this
    prompt -- the block's cross-attention K / V are then computed once and reused by every later call."""
    dt, dev = x.dtype, x.device
    n, d = x.shape
    sa, ca = block.self_attn, block.cross_attn
    heads = sa.num_heads
    if getattr(ca, "has_image_input", False):
        raise NotImplementedError("image-conditioned cross attention (k_img/v_img) is not on the wvd path")
    eps = _unwrap(block.norm1).eps
    mod = block_modulation(block, t_mod)
    # ---- self attention: x += gate_msa * o(attn(rope(rms(q)), rope(rms(k)), v)) ----
    h = ops.ln_modulate(x, mod[0], mod[1], eps=eps, out=ws.get("h", (n, d)))
    qkv = ws.get("qkv", (n, 3 * d))
    wb = [_lin(proj, dt, dev) for proj in (sa.q, sa.k, sa.v)]
    if exchange.world > 1 and hasattr(exchange, "v_ready"):
        # Ulysses: v first, so that its scatter over NVLink (which needs nothing from the RMSNorm / RoPE of q and k) runs
        # on a side stream under the q | k projection instead of after it
        ops.linear(h, wb[2][0], wb[2][1], out=qkv[:, 2 * d:])
        exchange.v_ready(ops, qkv, heads)
        ops.linear_grouped(h, [wb[0][0], wb[1][0]], [wb[0][1], wb[1][1]], out=qkv[:, :2 * d])
    else:
        ops.linear_grouped(h, [w for w, _ in wb], [b for _, b in wb], out=qkv)    # one launch for q | k | v
    a = exchange.norm_rope_attend(ops, qkv, heads, _norm_w(sa.norm_q, dt, dev), _norm_w(sa.norm_k, dt, dev),
                                  _unwrap(sa.norm_q).eps, rope, ws.get("attn", (n, d)), ws)
    w, b = _lin(sa.o, dt, dev)
    ops.linear(a, w, b, ops.EPI_BIAS_GATE_RES, gate=mod[2], residual=x, out=x)
    # ---- cross attention (ungated): x += o(attn(rms(q(norm3(x))), rms(k(ctx)), v(ctx))) ----
    n3 = _unwrap(block.norm3)
    h = ops.ln_modulate(x, weight=_param(n3.weight, dt, dev), bias=_param(n3.bias, dt, dev), eps=n3.eps,
                        out=ws.get("h", (n, d)))
    w, b = _lin(ca.q, dt, dev)
    q = ops.linear(h, w, b, out=ws.get("qc", (n, d)))
    ops.qk_rmsnorm_rope(q, None, _norm_w(ca.norm_q, dt, dev), None, _unwrap(ca.norm_q).eps)
    lc = context.shape[0]
    cached = None
    if text_entry is not None:
        stamp = TextCache.stamp(_unwrap(ca.k).weight, _unwrap(ca.v).weight, _unwrap(ca.norm_k).weight)
        cached = text_entry["kv"].get(id(block))
        if cached is not None and cached[2] != stamp:
            cached = None
    if cached is not None:
        kc, vc = cached[0], cached[1]
    else:
        own = text_entry is not None                     # cached K / V live in their own buffers, not in the workspace
        # k | v projections of the text in ONE launch (two M = 512 GEMMs of 80 tiles each leave half the SMs idle)
        kv = torch.empty((lc, 2 * d), dtype=dt, device=dev) if own else ws.get("kvc", (lc, 2 * d))
        (wk, bk), (wv, bv) = _lin(ca.k, dt, dev), _lin(ca.v, dt, dev)
        ops.linear_grouped(context, [wk, wv], [bk, bv], out=kv)
        kc, vc = kv[:, :d], kv[:, d:]
        ops.qk_rmsnorm_rope(kc, None, _norm_w(ca.norm_k, dt, dev), None, _unwrap(ca.norm_k).eps)
        if own:
            text_entry["kv"][id(block)] = (kc, vc, stamp)
    a = ops.attention(q, kc, vc, heads, out=ws.get("attn", (n, d)))
    w, b = _lin(ca.o, dt, dev)
    ops.linear(a, w, b, ops.EPI_BIAS_RES, residual=x, out=x)
    # ---- ffn: x += gate_mlp * W2 gelu_tanh(W1 modulate(norm2(x))) ----
    h = ops.ln_modulate(x, mod[3], mod[4], eps=_unwrap(block.norm2).eps, out=ws.get("h", (n, d)))
    w, b = _lin(block.ffn[0], dt, dev)
    mid = ops.linear(h, w, b, ops.EPI_BIAS_GELU, out=ws.get("mid", (n, w.shape[0])))
    w, b = _lin(block.ffn[2], dt, dev)
    ops.linear(mid, w, b, ops.EPI_BIAS_GATE_RES, gate=mod[5], residual=x, out=x)
    return x


def head_forward(head, x: Tensor, t: Tensor, ws: Workspace, ops=_cuda_ops) -> Tensor:
    """Head.forward (wan_video_dit.py:262-269): Linear(LN(x)*(1+scale)+shift), modulation + t (not t_mod)."""
    if t.dim() != 2 or t.shape[0] != 1:
        raise NotImplementedError("per-token head modulation (seperated_timestep) is not supported by the wvd path")
    dt, dev = x.dtype, x.device
    n, d = x.shape
    m = (head.modulation.to(dtype=t.dtype, device=t.device) + t.unsqueeze(1))[0].contiguous()   # (2, D): shift, scale
    h = ops.ln_modulate(x, m[0], m[1], eps=_unwrap(head.norm).eps, out=ws.get("h", (n, d)))
    w, b = _lin(head.head, dt, dev)
    return ops.linear(h, w, b, out=ws.get("head_out", (n, w.shape[0])))


def vace_forward(vace, x: Tensor, vace_context: Tensor, context: Tensor, t_mod: Tensor, rope: RopeInfo,
                 ws: Workspace, ops=_cuda_ops, exchange: SelfAttnExchange = _LOCAL,
                 token_slice: Optional[slice] = None, text_entry: Optional[dict] = None) -> Tensor:
    """VaceWanModel.forward + VaceWanAttentionBlock.forward (wan_video_vace.py:13-24, 53-87).

    Returns the hints as ONE preallocated (n_hints, tokens, D) buffer (the reference's O(k^2) stack/unbind copying
    disappears): c0 = before_proj(patch(vc)) + x ; c_{k+1} = block_k(c_k) ; hint_k = after_proj_k(c_{k+1}).
    x: (tokens, D) patch-embedded main stream (read only).  token_slice selects this rank's tokens (Ulysses)."""
    dt, dev = x.dtype, x.device
    n, d = x.shape
    if vace_context.shape[0] != 1:
        raise NotImplementedError("batched vace_context: loop over the batch in the caller")
    c = patch_embed(vace.vace_patch_embedding, vace_context.to(dt), ops)    # (tokens_total, D)
    if token_slice is not None:
        c = c[token_slice]
    if c.shape[0] > n:
        raise ValueError("vace_context has more tokens than the latent grid")
    c_buf = ws.get("vace_c", (n, d))
    c_buf[:c.shape[0]].copy_(c)
    if c.shape[0] < n:                                                      # zero-pad to len(x) (wan_video_vace.py:60-63)
        c_buf[c.shape[0]:].zero_()
    blocks = vace.vace_blocks
    hints = ws.get("vace_hints", (len(blocks), n, d))
    for j, blk in enumerate(blocks):
        if hasattr(blk, "before_proj"):
            w, b = _lin(blk.before_proj, dt, dev)
            c_new = ops.linear(c_buf, w, b, ops.EPI_BIAS_RES, residual=x, out=ws.get("vace_c2", (n, d)))
            c_buf = c_new
        dit_block_forward(blk, c_buf, context, t_mod, rope, ws, ops, exchange, text_entry)
        w, b = _lin(blk.after_proj, dt, dev)
        ops.linear(c_buf, w, b, out=hints[j])
    return hints
