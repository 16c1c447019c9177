"""Wan2.1 video DiT with the reference's module tree and state-dict keys, computing on libwvd.so (sm_100a).

Mirror of the operator/plugin interface of ``diffsynth/models/wan_video_dit.py`` (same class names, constructor
arguments, attribute names and ``state_dict()`` keys: ``blocks.i.self_attn.q.weight``, ``...norm_q.weight``,
``...modulation``, ``head.head.weight`` ...), so checkpoints, ``GeneralLoRALoader`` and ``enable_vram_management``
work unchanged.  Only the forwards differ: they call the C-ABI kernels through ``engine`` -- there is no PyTorch
fallback for the block math.  ``WanModel.forward`` is the *working* version of the reference's stale forward
(SURVEY.md 0.2): it equals ``model_fn_wan_video(dit=self, latents=x, ...)``.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import engine, ops

Tensor = torch.Tensor


def sinusoidal_embedding_1d(dim: int, position: Tensor) -> Tensor:
    """fp64 sinusoid [cos | sin], cast to position.dtype (wan_video_dit.py:68-72)."""
    half = dim // 2
    inv = torch.pow(10000, -torch.arange(half, dtype=torch.float64, device=position.device).div(half))
    ang = torch.outer(position.to(torch.float64), inv)
    return torch.cat([ang.cos(), ang.sin()], dim=1).to(position.dtype)


def precompute_freqs_cis(dim: int, end: int = 1024, theta: float = 10000.0) -> Tensor:
    """1-D RoPE table, complex128 (end, dim//2) (wan_video_dit.py:83-89)."""
    inv = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].double() / dim))
    ang = torch.outer(torch.arange(end, device=inv.device), inv)
    return torch.polar(torch.ones_like(ang), ang)


def precompute_freqs_cis_3d(dim: int, end: int = 1024, theta: float = 10000.0):
    """(frame, height, width) tables with dim - 2*(dim//3), dim//3, dim//3 channels (wan_video_dit.py:75-80)."""
    return (precompute_freqs_cis(dim - 2 * (dim // 3), end, theta), precompute_freqs_cis(dim // 3, end, theta),
            precompute_freqs_cis(dim // 3, end, theta))


def modulate(x: Tensor, shift: Tensor, scale: Tensor) -> Tensor:
    """LN-free form kept for API compatibility (wan_video_dit.py:64-65); the fused kernel is ops.ln_modulate."""
    return x * (1 + scale) + shift


def flash_attention(q: Tensor, k: Tensor, v: Tensor, num_heads: int, compatibility_mode: bool = False) -> Tensor:
    """'b s (n d)' attention on the tcgen05 kernel (replaces the FA3/FA2/Sage/SDPA dispatch, wan_video_dit.py:28-61)."""
    outs = [ops.attention(q[b], k[b], v[b], num_heads) for b in range(q.shape[0])]
    return torch.stack(outs, dim=0)


class RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x: Tensor) -> Tensor:
        """norm(x.float()).to(dtype) * weight (wan_video_dit.py:106-111); x is NOT modified."""
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        y = torch.empty_like(x2)
        ops.qk_rmsnorm_rope(x2, None, self.weight.to(x.dtype), None, self.eps, q_out=y)
        return y.view(x.shape)


class AttentionModule(nn.Module):
    def __init__(self, num_heads: int):
        super().__init__()
        self.num_heads = num_heads

    def forward(self, q, k, v):
        return flash_attention(q, k, v, self.num_heads)


class SelfAttention(nn.Module):
    def __init__(self, dim: int, num_heads: int, eps: float = 1e-6):
        super().__init__()
        self.dim, self.num_heads, self.head_dim = dim, num_heads, dim // num_heads
        self.q, self.k, self.v, self.o = (nn.Linear(dim, dim) for _ in range(4))
        self.norm_q, self.norm_k = RMSNorm(dim, eps=eps), RMSNorm(dim, eps=eps)
        self.attn = AttentionModule(num_heads)


class CrossAttention(nn.Module):
    def __init__(self, dim: int, num_heads: int, eps: float = 1e-6, has_image_input: bool = False):
        super().__init__()
        if has_image_input:
            raise NotImplementedError("has_image_input (I2V CLIP branch) is outside the T2V/VACE hot path")
        self.dim, self.num_heads, self.head_dim = dim, num_heads, dim // num_heads
        self.q, self.k, self.v, self.o = (nn.Linear(dim, dim) for _ in range(4))
        self.norm_q, self.norm_k = RMSNorm(dim, eps=eps), RMSNorm(dim, eps=eps)
        self.has_image_input = has_image_input
        self.attn = AttentionModule(num_heads)


class GateModule(nn.Module):
    def forward(self, x, gate, residual):
        n, d = x.shape[-2:]
        return ops.gate_residual(x.reshape(-1, d).contiguous(), gate.reshape(-1), residual.reshape(-1, d).contiguous()).view_as(x)


class DiTBlock(nn.Module):
    def __init__(self, has_image_input: bool, dim: int, num_heads: int, ffn_dim: int, eps: float = 1e-6):
        super().__init__()
        self.dim, self.num_heads, self.ffn_dim = dim, num_heads, ffn_dim
        self.self_attn = SelfAttention(dim, num_heads, eps)
        self.cross_attn = CrossAttention(dim, num_heads, eps, has_image_input=has_image_input)
        self.norm1 = nn.LayerNorm(dim, eps=eps, elementwise_affine=False)
        self.norm2 = nn.LayerNorm(dim, eps=eps, elementwise_affine=False)
        self.norm3 = nn.LayerNorm(dim, eps=eps)
        self.ffn = nn.Sequential(nn.Linear(dim, ffn_dim), nn.GELU(approximate="tanh"), nn.Linear(ffn_dim, dim))
        self.modulation = nn.Parameter(torch.randn(1, 6, dim) / dim ** 0.5)
        self.gate = GateModule()

    def forward(self, x: Tensor, context: Tensor, t_mod: Tensor, freqs) -> Tensor:
        """block(x, context, t_mod, freqs) -> x  (wan_video_dit.py:214-230).  ``freqs`` is the reference's complex
        (N, 1, 64) tensor (wan_video_new.py:1392-1396) or an engine.RopeInfo (model_fn / WanModel.rope_info: the
        device-resident table, no per-call gather); x (1, N, D) is NOT modified (the engine works on a copy)."""
        x2 = ops.as_2d(x).clone()
        ctx = ops.as_2d(context)
        rope = engine.as_rope_info(freqs, x2.device)
        ws = engine.workspace(x2.shape[0], self.dim, self.ffn_dim, ctx.shape[0], x2.dtype, x2.device)
        return engine.dit_block_forward(self, x2, ctx, t_mod, rope, ws).view_as(x)


class Head(nn.Module):
    def __init__(self, dim: int, out_dim: int, patch_size: Tuple[int, int, int], eps: float):
        super().__init__()
        self.dim, self.patch_size = dim, patch_size
        self.norm = nn.LayerNorm(dim, eps=eps, elementwise_affine=False)
        self.head = nn.Linear(dim, out_dim * math.prod(patch_size))
        self.modulation = nn.Parameter(torch.randn(1, 2, dim) / dim ** 0.5)

    def forward(self, x: Tensor, t_mod: Tensor) -> Tensor:
        x2 = ops.as_2d(x)
        ws = engine.workspace(x2.shape[0], self.dim, 0, 0, x2.dtype, x2.device)
        return engine.head_forward(self, x2, t_mod, ws).clone().unsqueeze(0)


class WanModel(nn.Module):
    """Same constructor as the reference (wan_video_dit.py:273-294); T2V / VACE feature set."""

    def __init__(self, dim: int, in_dim: int, ffn_dim: int, out_dim: int, text_dim: int, freq_dim: int, eps: float,
                 patch_size: Tuple[int, int, int], num_heads: int, num_layers: int, has_image_input: bool = False,
                 has_image_pos_emb: bool = False, has_ref_conv: bool = False, add_control_adapter: bool = False,
                 in_dim_control_adapter: int = 24, seperated_timestep: bool = False, require_vae_embedding: bool = True,
                 require_clip_embedding: bool = True, fuse_vae_embedding_in_latents: bool = False):
        super().__init__()
        if has_image_input or add_control_adapter or seperated_timestep:
            raise NotImplementedError("I2V / camera-control / TI2V variants are outside the T2V+VACE hot path")
        if dim // num_heads != 128:
            raise ValueError("the wvd attention/RoPE kernels are specialised for head_dim 128 (every Wan model)")
        self.dim, self.in_dim, self.freq_dim = dim, in_dim, freq_dim
        self.has_image_input, self.patch_size = has_image_input, tuple(patch_size)
        self.seperated_timestep = seperated_timestep
        self.require_vae_embedding, self.require_clip_embedding = require_vae_embedding, require_clip_embedding
        self.fuse_vae_embedding_in_latents = fuse_vae_embedding_in_latents
        self.has_image_pos_emb, self.has_ref_conv = has_image_pos_emb, has_ref_conv
        self.patch_embedding = nn.Conv3d(in_dim, dim, kernel_size=patch_size, stride=patch_size)
        self.text_embedding = nn.Sequential(nn.Linear(text_dim, dim), nn.GELU(approximate="tanh"), nn.Linear(dim, dim))
        self.time_embedding = nn.Sequential(nn.Linear(freq_dim, dim), nn.SiLU(), nn.Linear(dim, dim))
        self.time_projection = nn.Sequential(nn.SiLU(), nn.Linear(dim, dim * 6))
        self.blocks = nn.ModuleList([DiTBlock(has_image_input, dim, num_heads, ffn_dim, eps) for _ in range(num_layers)])
        self.head = Head(dim, out_dim, patch_size, eps)
        self.freqs = precompute_freqs_cis_3d(dim // num_heads)
        if has_ref_conv:
            self.ref_conv = nn.Conv2d(16, dim, kernel_size=(2, 2), stride=(2, 2))
        self.control_adapter = None
        self._rope_table = None

    # -- reference helper surface (wan_video_dit.py:339-352) --
    def patchify(self, x: Tensor, control_camera_latents_input: Optional[Tensor] = None) -> Tensor:
        if control_camera_latents_input is not None:
            raise NotImplementedError("camera control is outside the T2V+VACE hot path")
        return self.patch_embedding(x)

    def unpatchify(self, x: Tensor, grid_size) -> Tensor:
        f, h, w = (int(g) for g in grid_size)
        pt, ph, pw = self.patch_size
        b = x.shape[0]
        return (x.view(b, f, h, w, pt, ph, pw, -1).permute(0, 7, 1, 4, 2, 5, 3, 6)
                .reshape(b, -1, f * pt, h * ph, w * pw))

    def rope_info(self, f: int, h: int, w: int, device, rope_indices: Optional[Tensor] = None,
                  token_offset: int = 0) -> engine.RopeInfo:
        """Device-resident cos/sin table built once from ``self.freqs`` (replaces the per-call 30 MB complex128
        gather + H2D copy of wan_video_new.py:1392-1396)."""
        if self._rope_table is None or self._rope_table.device != torch.device(device):
            self._rope_table = ops.make_rope_table(self.freqs, device)
        fid = None
        if rope_indices is not None:
            fid = torch.as_tensor(rope_indices, dtype=torch.int32).to(device).contiguous()
            f = int(fid.numel())
        return engine.RopeInfo(self._rope_table, (f, h, w), token_offset, fid)

    def forward(self, x: Tensor, timestep: Tensor, context: Tensor, clip_feature: Optional[Tensor] = None,
                y: Optional[Tensor] = None, rope_indices: Optional[Tensor] = None,
                use_gradient_checkpointing: bool = False, use_gradient_checkpointing_offload: bool = False, **kwargs):
        from .pipeline import model_fn_wan_video
        return model_fn_wan_video(dit=self, latents=x, timestep=timestep, context=context, clip_feature=clip_feature,
                                  y=y, rope_indices=rope_indices, **kwargs)

    @staticmethod
    def state_dict_converter():
        return WanModelStateDictConverter()


class WanModelStateDictConverter:
    """Shape-driven replacement of the reference's hash tables (wan_video_dit.py:506-751): the config is inferred
    from tensor shapes, so any T2V / VACE checkpoint in the native ("civitai") key format loads."""

    def from_civitai(self, state_dict):
        sd = {}
        for k, v in state_dict.items():
            if k.startswith("vace"):
                continue                       # VACE weights go to VaceWanModel (wan_video_vace.py:98-110)
            if k.startswith("model.diffusion_model."):
                k = k[len("model.diffusion_model."):]
            sd[k] = v
        return sd, infer_dit_config(sd)

    def from_diffusers(self, state_dict):
        raise NotImplementedError("diffusers-format Wan checkpoints: convert to the native key format first")


def infer_dit_config(sd) -> dict:
    pe = sd["patch_embedding.weight"]
    dim, in_dim = int(pe.shape[0]), int(pe.shape[1])
    layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    out_ch = int(sd["head.head.weight"].shape[0]) // math.prod(pe.shape[2:])
    return dict(dim=dim, in_dim=in_dim, ffn_dim=int(sd["blocks.0.ffn.0.weight"].shape[0]), out_dim=out_ch,
                text_dim=int(sd["text_embedding.0.weight"].shape[1]), freq_dim=int(sd["time_embedding.0.weight"].shape[1]),
                eps=1e-6, patch_size=tuple(int(s) for s in pe.shape[2:]), num_heads=dim // 128, num_layers=layers,
                has_image_input="img_emb.proj.0.weight" in sd, has_ref_conv="ref_conv.weight" in sd)
