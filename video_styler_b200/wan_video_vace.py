"""VACE hint model with the reference's module tree / state-dict keys (diffsynth/models/wan_video_vace.py),
computing on libwvd.so.  The Ditto LoRA is merged into these weights at load time by the reference's
``GeneralLoRALoader`` (``named_modules()`` yields ``vace_blocks.j.self_attn.q`` ... exactly as upstream)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import engine, ops
from .wan_video_dit import DiTBlock

Tensor = torch.Tensor


class VaceWanAttentionBlock(DiTBlock):
    def __init__(self, has_image_input, dim, num_heads, ffn_dim, eps=1e-6, block_id=0):
        super().__init__(has_image_input, dim, num_heads, ffn_dim, eps=eps)
        self.block_id = block_id
        if block_id == 0:
            self.before_proj = nn.Linear(dim, dim)
        self.after_proj = nn.Linear(dim, dim)

    def forward(self, c, x, context, t_mod, freqs):
        """The reference's stacked-tensor protocol (wan_video_vace.py:13-24) over the engine, for callers that drive
        the blocks themselves: block 0 takes c (1, N, D) and returns stack([skip_0, c_1]); block k takes
        stack([skip_0 .. skip_{k-1}, c_k]) and returns stack([skip_0 .. skip_k, c_{k+1}]).
        ``VaceWanModel.forward`` does not go through here: it writes the hints once into a preallocated buffer."""
        if self.block_id == 0:
            c2 = ops.as_2d(c)
            w, b = engine._lin(self.before_proj, c2.dtype, c2.device)
            cur = ops.linear(c2.contiguous(), w, b, ops.EPI_BIAS_RES, residual=ops.as_2d(x).contiguous())   # before_proj(c) + x
            skips = []
        else:
            skips = list(torch.unbind(c))
            cur = ops.as_2d(skips.pop(-1)).clone()
        ctx = ops.as_2d(context)
        rope = engine.as_rope_info(freqs, cur.device)
        ws = engine.workspace(cur.shape[0], self.dim, self.ffn_dim, ctx.shape[0], cur.dtype, cur.device)
        engine.dit_block_forward(self, cur, ctx, t_mod, rope, ws)                     # DiTBlock.forward, in place on cur
        w, b = engine._lin(self.after_proj, cur.dtype, cur.device)
        skip = ops.linear(cur, w, b)                                                   # c_skip = after_proj(c)
        like = skips[0] if skips else c
        return torch.stack(skips + [skip.view(like.shape), cur.view(like.shape)])


class VaceWanModel(nn.Module):
    def __init__(self, vace_layers=(0, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28), vace_in_dim=96,
                 patch_size=(1, 2, 2), has_image_input=False, dim=1536, num_heads=12, ffn_dim=8960, eps=1e-6):
        super().__init__()
        self.vace_layers = tuple(vace_layers)
        self.vace_in_dim = vace_in_dim
        self.vace_layers_mapping = {i: n for n, i in enumerate(self.vace_layers)}
        self.vace_blocks = nn.ModuleList([VaceWanAttentionBlock(has_image_input, dim, num_heads, ffn_dim, eps, block_id=i)
                                          for i in self.vace_layers])
        self.vace_patch_embedding = nn.Conv3d(vace_in_dim, dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x: Tensor, vace_context: Tensor, context: Tensor, t_mod: Tensor, freqs,
                use_gradient_checkpointing: bool = False, use_gradient_checkpointing_offload: bool = False):
        """vace(x, vace_context, context, t_mod, freqs) -> tuple of (1, N, D) hints (wan_video_vace.py:53-87).
        ``freqs`` is an engine.RopeInfo or the reference's complex (N, 1, 64) tensor.  The hints are views of one
        workspace buffer (valid until the next call)."""
        x2, ctx = ops.as_2d(x), ops.as_2d(context)
        freqs = engine.as_rope_info(freqs, x2.device)
        blk = self.vace_blocks[0]
        ws = engine.workspace(x2.shape[0], blk.dim, blk.ffn_dim, ctx.shape[0], x2.dtype, x2.device)
        hints = engine.vace_forward(self, x2, vace_context, ctx, t_mod, freqs, ws)
        return tuple(h.unsqueeze(0) for h in hints.unbind(0))

    @staticmethod
    def state_dict_converter():
        return VaceWanModelDictConverter()


class VaceWanModelDictConverter:
    """Keeps the ``vace*`` keys and infers the config from shapes (the reference hashes key names, wan_video_vace.py:98-113)."""

    def from_civitai(self, state_dict):
        sd = {k: v for k, v in state_dict.items() if k.startswith("vace")}
        if not sd:
            return sd, {}
        pe = sd["vace_patch_embedding.weight"]
        dim = int(pe.shape[0])
        n_blocks = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("vace_blocks."))
        layers = {8: (0, 5, 10, 15, 20, 25, 30, 35), 15: tuple(range(0, 30, 2))}.get(n_blocks)
        if layers is None:
            raise ValueError(f"cannot infer vace_layers for {n_blocks} VACE blocks")
        cfg = dict(vace_layers=layers, vace_in_dim=int(pe.shape[1]), patch_size=tuple(int(s) for s in pe.shape[2:]),
                   has_image_input=False, dim=dim, num_heads=dim // 128,
                   ffn_dim=int(sd["vace_blocks.0.ffn.0.weight"].shape[0]), eps=1e-6)
        return sd, cfg
