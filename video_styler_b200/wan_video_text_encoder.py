"""umT5-XXL text encoder on the B200 path.

Mirrors ``diffsynth/models/wan_video_text_encoder.py`` -- same class names, constructor arguments, attribute names and
``state_dict()`` keys (``token_embedding.weight``, ``blocks.i.norm1.weight``, ``blocks.i.attn.{q,k,v,o}.weight``,
``blocks.i.pos_embedding.embedding.weight``, ``blocks.i.ffn.{gate.0,fc1,fc2}.weight``, ``norm.weight``), so the
reference's checkpoints load unchanged -- with ``WanTextEncoder.forward`` on libwvd.so:

    T5LayerNorm (RMS over the full width)      wvd_qk_rmsnorm_rope without RoPE (same rounding points: fp32 normalise,
                                               cast, * weight)                                 ref :22-35
    q | k | v projections (no bias)            one grouped tcgen05 GEMM launch                 ref :47-50, 64-66
    attention with relative-position bias,     wvd_attention_bias_fwd (head_dim 64, no scaling, fp32 softmax, the
    key mask                                   reference's bf16 rounding points)               ref :68-84, 141-175
    o projection + residual                    GEMM, residual epilogue                         ref :87-88, 136
    gated-GELU feed-forward                    gate GEMM whose epilogue is the encoder's hand-written GELU with its seven
                                               per-op bf16 roundings (WVD_EPI_BIAS_GELU_T5, ref :16-20), fc1 GEMM with the
                                               elementwise-product epilogue (WVD_EPI_BIAS_MUL), fc2 GEMM with the
                                               residual epilogue                               ref :93-113, 137
The token-embedding gather and the 1,023-entry bucket table stay in PyTorch (index glue, no arithmetic).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import engine, ops as _cuda_ops

Tensor = torch.Tensor


class GELU(nn.Module):
    """Parameter-free placeholder so that ``ffn.gate`` keeps the reference's Sequential(Linear, GELU) key layout."""

    def forward(self, x):
        return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


class T5LayerNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.dim, self.eps = dim, eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x: Tensor, ops=_cuda_ops) -> Tensor:
        x2 = x.reshape(-1, self.dim)
        out = torch.empty_like(x2)
        ops.qk_rmsnorm_rope(x2, None, self.weight.to(x.dtype), None, self.eps, q_out=out)
        return out.view_as(x)


class T5Attention(nn.Module):
    def __init__(self, dim: int, dim_attn: int, num_heads: int, dropout: float = 0.1):
        super().__init__()
        assert dim_attn % num_heads == 0
        self.dim, self.dim_attn, self.num_heads, self.head_dim = dim, dim_attn, num_heads, dim_attn // num_heads
        self.q = nn.Linear(dim, dim_attn, bias=False)
        self.k = nn.Linear(dim, dim_attn, bias=False)
        self.v = nn.Linear(dim, dim_attn, bias=False)
        self.o = nn.Linear(dim_attn, dim, bias=False)
        self.dropout = nn.Dropout(dropout)


class T5FeedForward(nn.Module):
    def __init__(self, dim: int, dim_ffn: int, dropout: float = 0.1):
        super().__init__()
        self.dim, self.dim_ffn = dim, dim_ffn
        self.gate = nn.Sequential(nn.Linear(dim, dim_ffn, bias=False), GELU())
        self.fc1 = nn.Linear(dim, dim_ffn, bias=False)
        self.fc2 = nn.Linear(dim_ffn, dim, bias=False)
        self.dropout = nn.Dropout(dropout)


def relative_position_bucket(rel: Tensor, num_buckets: int, max_dist: int = 128, bidirectional: bool = True) -> Tensor:
    """T5 bucket of a key-minus-query offset (ref :155-175).  Half of the buckets per direction when bidirectional; in a
    direction the first half of the buckets holds the exact distances 0 .. n/2 - 1, the second half log-spaced distances
    up to ``max_dist`` (everything beyond shares the last bucket).  Same float32 log / truncation as the reference."""
    if bidirectional:
        per_dir = num_buckets // 2
        base = torch.where(rel > 0, per_dir, 0)
        dist = rel.abs()
    else:
        per_dir = num_buckets
        base = torch.zeros_like(rel)
        dist = (-rel).clamp(min=0)
    exact = per_dir // 2
    log_part = torch.log(dist.float() / exact) / math.log(max_dist / exact) * (per_dir - exact)
    far = (exact + log_part.long()).clamp(max=per_dir - 1)
    return base + torch.where(dist < exact, dist, far)


_BUCKET_CACHE = {}


class T5RelativeEmbedding(nn.Module):
    def __init__(self, num_buckets: int, num_heads: int, bidirectional: bool, max_dist: int = 128):
        super().__init__()
        self.num_buckets, self.num_heads, self.bidirectional, self.max_dist = num_buckets, num_heads, bidirectional, max_dist
        self.embedding = nn.Embedding(num_buckets, num_heads)

    def buckets(self, lq: int, lk: int, device) -> Tensor:
        """bucket(j - i) for every offset -(lq-1) .. lk-1: index glue, the same for every layer and every call."""
        key = (lq, lk, str(device))
        hit = _BUCKET_CACHE.get((key, self.num_buckets, self.max_dist, self.bidirectional))
        if hit is None:
            rel = torch.arange(-(lq - 1), lk, device=device)
            hit = relative_position_bucket(rel, self.num_buckets, self.max_dist, self.bidirectional)
            if len(_BUCKET_CACHE) > 16:
                _BUCKET_CACHE.clear()
            _BUCKET_CACHE[(key, self.num_buckets, self.max_dist, self.bidirectional)] = hit
        return hit

    def table(self, lq: int, lk: int, dtype, device) -> Tensor:
        """The bias as the kernel reads it: (heads, lq + lk - 1), entry [h][(j - i) + lq - 1] = embedding[bucket(j - i)][h]
        -- the (1, N, Lq, Lk) tensor of the reference's forward (:141-153) depends on j - i only."""
        emb = self.embedding.weight
        if emb.dtype != dtype or emb.device != torch.device(device):
            emb = emb.to(device=device, dtype=dtype)
        return emb[self.buckets(lq, lk, device)].t().contiguous()

    def forward(self, lq: int, lk: int) -> Tensor:
        """The dense (1, N, Lq, Lk) bias the reference module returns, expanded from the table."""
        tab = self.table(lq, lk, self.embedding.weight.dtype, self.embedding.weight.device)      # (N, lq + lk - 1)
        idx = (torch.arange(lk, device=tab.device)[None, :] - torch.arange(lq, device=tab.device)[:, None]) + (lq - 1)
        return tab[:, idx].unsqueeze(0).contiguous()


class T5SelfAttention(nn.Module):
    def __init__(self, dim, dim_attn, dim_ffn, num_heads, num_buckets, shared_pos=True, dropout=0.1):
        super().__init__()
        self.dim, self.dim_attn, self.dim_ffn, self.num_heads = dim, dim_attn, dim_ffn, num_heads
        self.num_buckets, self.shared_pos = num_buckets, shared_pos
        self.norm1 = T5LayerNorm(dim)
        self.attn = T5Attention(dim, dim_attn, num_heads, dropout)
        self.norm2 = T5LayerNorm(dim)
        self.ffn = T5FeedForward(dim, dim_ffn, dropout)
        self.pos_embedding = None if shared_pos else T5RelativeEmbedding(num_buckets, num_heads, bidirectional=True)


def init_weights(m):
    """The reference's initialisation (ref :177-194): normal weights with a per-layer-kind standard deviation."""
    std = {}
    if isinstance(m, T5LayerNorm):
        nn.init.ones_(m.weight)
    elif isinstance(m, T5FeedForward):
        std = {m.gate[0]: m.dim ** -0.5, m.fc1: m.dim ** -0.5, m.fc2: m.dim_ffn ** -0.5}
    elif isinstance(m, T5Attention):
        std = {m.q: (m.dim * m.dim_attn) ** -0.5, m.k: m.dim ** -0.5, m.v: m.dim ** -0.5,
               m.o: (m.num_heads * m.dim_attn) ** -0.5}
    elif isinstance(m, T5RelativeEmbedding):
        std = {m.embedding: (2 * m.num_buckets * m.num_heads) ** -0.5}
    for layer, sd in std.items():
        nn.init.normal_(layer.weight, std=sd)


def _w(lin, dtype, device) -> Tensor:
    w = engine._unwrap(lin).weight
    return w if (w.dtype == dtype and w.device == device) else w.to(device=device, dtype=dtype)


def t5_block_forward(block: T5SelfAttention, x: Tensor, key_mask: Optional[Tensor], bias: Optional[Tensor], ws: dict,
                     ops=_cuda_ops) -> Tensor:
    """T5SelfAttention.forward (ref :130-138) on a (tokens, dim) stream, updated IN PLACE."""
    dt, dev = x.dtype, x.device
    n, d = x.shape
    at, ff = block.attn, block.ffn
    da = at.dim_attn

    def buf(name, shape):
        t = ws.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dt:
            t = ws[name] = torch.empty(shape, dtype=dt, device=dev)
        return t

    h = buf("h", (n, d))
    ops.qk_rmsnorm_rope(x, None, _w(block.norm1, dt, dev), None, engine._unwrap(block.norm1).eps, q_out=h)
    qkv = buf("qkv", (n, 3 * da))
    ops.linear_grouped(h, [_w(at.q, dt, dev), _w(at.k, dt, dev), _w(at.v, dt, dev)], [None, None, None], out=qkv)
    a = ops.attention_bias(qkv[:, :da], qkv[:, da:2 * da], qkv[:, 2 * da:], at.num_heads, bias, key_mask, 1.0,
                           out=buf("attn", (n, da)))
    ops.linear(a, _w(at.o, dt, dev), None, ops.EPI_BIAS_RES, residual=x, out=x)                 # x + attn(norm1(x))
    ops.qk_rmsnorm_rope(x, None, _w(block.norm2, dt, dev), None, engine._unwrap(block.norm2).eps, q_out=h)
    g = ops.linear(h, _w(ff.gate[0], dt, dev), None, ops.EPI_BIAS_GELU_T5, out=buf("gate", (n, ff.dim_ffn)))
    u = ops.linear(h, _w(ff.fc1, dt, dev), None, ops.EPI_BIAS_MUL, residual=g, out=buf("mid", (n, ff.dim_ffn)))
    ops.linear(u, _w(ff.fc2, dt, dev), None, ops.EPI_BIAS_RES, residual=x, out=x)                # x + ffn(norm2(x))
    return x


class WanTextEncoder(nn.Module):
    """Same constructor as the reference (ref :197-231); forward(ids, mask) -> (B, L, dim)."""

    def __init__(self, vocab=256384, dim=4096, dim_attn=4096, dim_ffn=10240, num_heads=64, num_layers=24, num_buckets=32,
                 shared_pos=False, dropout=0.1):
        super().__init__()
        if dim_attn // num_heads != 64:
            raise ValueError("the wvd umT5 attention kernel is specialised for head_dim 64 (umT5-XXL: 4096 / 64 heads)")
        self.dim, self.dim_attn, self.dim_ffn, self.num_heads = dim, dim_attn, dim_ffn, num_heads
        self.num_layers, self.num_buckets, self.shared_pos = num_layers, num_buckets, shared_pos
        self.token_embedding = vocab if isinstance(vocab, nn.Embedding) else nn.Embedding(vocab, dim)
        self.pos_embedding = T5RelativeEmbedding(num_buckets, num_heads, bidirectional=True) if shared_pos else None
        self.dropout = nn.Dropout(dropout)
        self.blocks = nn.ModuleList([T5SelfAttention(dim, dim_attn, dim_ffn, num_heads, num_buckets, shared_pos, dropout)
                                     for _ in range(num_layers)])
        self.norm = T5LayerNorm(dim)
        self.apply(init_weights)
        self._ws = {}

    def forward(self, ids: Tensor, mask: Optional[Tensor] = None, ops=_cuda_ops) -> Tensor:
        if self.training and self.dropout.p > 0:
            raise NotImplementedError("the wvd text encoder is inference-only: call .eval()")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("the wvd path is forward-only: call under torch.no_grad()")
        emb = engine._unwrap(self.token_embedding)
        dt, dev = emb.weight.dtype, ids.device
        b, l = ids.shape
        if mask is not None and mask.dim() != 2:
            raise NotImplementedError("(B, L1, L2) attention masks are not on the wvd path (the pipeline passes (B, L))")
        outs = []
        for i in range(b):
            table = emb.weight if emb.weight.device == dev else emb.weight.to(dev)      # offloaded by a vram wrapper
            x = torch.nn.functional.embedding(ids[i], table).to(dtype=dt).contiguous()  # (L, dim) gather
            km = None if mask is None else (mask[i] != 0).to(torch.int32).contiguous()
            shared = self.pos_embedding.table(l, l, dt, dev) if self.shared_pos else None
            for block in self.blocks:
                bias = shared if self.shared_pos else engine._unwrap(block.pos_embedding).table(l, l, dt, dev)
                t5_block_forward(block, x, km, bias, self._ws, ops)
            out = torch.empty_like(x)
            ops.qk_rmsnorm_rope(x, None, _w(self.norm, dt, dev), None, engine._unwrap(self.norm).eps, q_out=out)
            outs.append(out)
        return torch.stack(outs, dim=0)

    @staticmethod
    def state_dict_converter():
        return WanTextEncoderStateDictConverter()


class WanTextEncoderStateDictConverter:
    def from_diffusers(self, state_dict):
        return state_dict

    def from_civitai(self, state_dict):
        return state_dict


def encode_prompt(text_encoder: WanTextEncoder, ids: Tensor, mask: Tensor) -> Tensor:
    """WanPrompter.encode_prompt after tokenisation (prompters/wan_prompter.py:102-109): encode, then zero the
    embedding rows past each prompt's length."""
    seq_lens = mask.gt(0).sum(dim=1).long()
    prompt_emb = text_encoder(ids, mask)
    for i, v in enumerate(seq_lens):
        prompt_emb[:, v:] = 0          # the reference's own slicing (all rows, wan_prompter.py:108); batch is 1 per prompt
    return prompt_emb
