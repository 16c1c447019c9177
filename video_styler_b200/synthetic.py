"""Synthetic (random-init) models and inputs of the BASELINE.json configs -- there are no checkpoints offline.

Weights follow nn.Linear/Conv3d default init (U(-1/sqrt(fan_in), 1/sqrt(fan_in))), modulation ~ N(0,1)/sqrt(D),
norm affine = (1, 0); they are drawn directly on the target device in the target dtype, so a
17-B-parameter Wan2.1-VACE-14B materialises in seconds.  The Ditto LoRA stand-in (rank 128 on q,k,v,o,ffn.0,ffn.2
of every VACE block, train.sh:16-18) is merged at load like GeneralLoRALoader does: W += alpha * (B @ A).
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from .wan_video_dit import WanModel, precompute_freqs_cis_3d
from .wan_video_vace import VaceWanModel

DIT_CONFIGS = {
    "1.3B": dict(dim=1536, in_dim=16, ffn_dim=8960, out_dim=16, text_dim=4096, freq_dim=256, eps=1e-6,
                 patch_size=(1, 2, 2), num_heads=12, num_layers=30),
    "14B": dict(dim=5120, in_dim=16, ffn_dim=13824, out_dim=16, text_dim=4096, freq_dim=256, eps=1e-6,
                patch_size=(1, 2, 2), num_heads=40, num_layers=40),
}
VACE_CONFIGS = {
    "1.3B": dict(vace_layers=tuple(range(0, 30, 2)), vace_in_dim=96, patch_size=(1, 2, 2), dim=1536, num_heads=12,
                 ffn_dim=8960, eps=1e-6),
    "14B": dict(vace_layers=(0, 5, 10, 15, 20, 25, 30, 35), vace_in_dim=96, patch_size=(1, 2, 2), dim=5120,
                num_heads=40, ffn_dim=13824, eps=1e-6),
}
LORA_TARGETS = ("self_attn.q", "self_attn.k", "self_attn.v", "self_attn.o", "cross_attn.q", "cross_attn.k",
                "cross_attn.v", "cross_attn.o", "ffn.0", "ffn.2")
# named workloads: latent (B, C, F, H, W); tokens = F * H/2 * W/2
WORKLOADS = {
    "c1": dict(size="1.3B", vace=False, latent=(1, 16, 5, 32, 32)),      # 1,280 tokens
    "c2": dict(size="1.3B", vace=False, latent=(1, 16, 21, 60, 104)),    # 32,760 tokens
    "c3": dict(size="14B", vace=True, latent=(1, 16, 19, 60, 104)),      # 29,640 tokens
    "c5": dict(size="14B", vace=True, latent=(1, 16, 21, 90, 160)),      # 75,600 tokens
}


def _materialize(module: torch.nn.Module, device, dtype, gen: torch.Generator):
    module.to(dtype=dtype)
    module.to_empty(device=device)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("modulation"):
                p.copy_(torch.randn(p.shape, device=device, generator=gen) / math.sqrt(p.shape[-1]))
            elif ".norm_q." in name or ".norm_k." in name or name.endswith("norm3.weight"):
                p.fill_(1.0)
            elif name.endswith("norm3.bias"):
                p.zero_()
            elif name.endswith(".weight"):
                fan_in = math.prod(p.shape[1:])
                p.uniform_(-1.0 / math.sqrt(fan_in), 1.0 / math.sqrt(fan_in), generator=gen)
            else:
                p.uniform_(-0.02, 0.02, generator=gen)
    return module.eval().requires_grad_(False)


def merge_lora_standin(vace: VaceWanModel, rank: int = 128, alpha: float = 1.0, seed: int = 2):
    """W <- W + alpha * (B @ A) in the weight dtype on the device (diffsynth/lora/__init__.py:28-45)."""
    mods = dict(vace.named_modules())
    dev = next(vace.parameters()).device
    gen = torch.Generator(device=dev).manual_seed(seed)
    n = 0
    with torch.no_grad():
        for j in range(len(vace.vace_blocks)):
            for tgt in LORA_TARGETS:
                lin = mods[f"vace_blocks.{j}.{tgt}"]
                out_f, in_f = lin.weight.shape
                a = (torch.randn(rank, in_f, device=dev, generator=gen) / math.sqrt(in_f)).to(lin.weight.dtype)
                b = (torch.randn(out_f, rank, device=dev, generator=gen) * 0.01).to(lin.weight.dtype)
                lin.weight.add_(alpha * torch.mm(b, a))
                n += 1
    return n


def build_models(size: str = "14B", with_vace: bool = True, device="cuda", dtype=torch.bfloat16, seed: int = 0,
                 lora_rank: Optional[int] = 128, num_layers: Optional[int] = None):
    cfg = dict(DIT_CONFIGS[size])
    if num_layers is not None:
        cfg["num_layers"] = num_layers
    gen = torch.Generator(device=device).manual_seed(seed)
    with torch.device("meta"):
        dit = WanModel(has_image_input=False, **cfg)
    dit = _materialize(dit, device, dtype, gen)
    dit.freqs = precompute_freqs_cis_3d(cfg["dim"] // cfg["num_heads"])
    vace = None
    if with_vace:
        vcfg = dict(VACE_CONFIGS[size])
        if num_layers is not None:
            vcfg["vace_layers"] = tuple(l for l in vcfg["vace_layers"] if l < num_layers)
        with torch.device("meta"):
            vace = VaceWanModel(has_image_input=False, **vcfg)
        vace = _materialize(vace, device, dtype, gen)
        if lora_rank:
            merge_lora_standin(vace, lora_rank)
    return dit, vace


def make_inputs(latent_shape, text_dim: int = 4096, with_vace: bool = True, seed: int = 1, dtype=torch.bfloat16,
                pin: bool = True, text_len: int = 512, prompt_len: int = 64):
    """HOST tensors (pinned): latents ~ N(0,1); context ~ N(0,1) with rows >= prompt_len zeroed
    (prompters/wan_prompter.py:107-108); vace_context channels 0-31 ~ N(0,1), 32-95 = 1 (wan_video_new.py:882-894)."""
    g = torch.Generator().manual_seed(seed)
    b, c, f, h, w = latent_shape
    lat = torch.randn(b, c, f, h, w, generator=g).to(dtype)
    ctx = torch.randn(b, text_len, text_dim, generator=g)
    ctx[:, prompt_len:] = 0
    out = dict(latents=lat, context=ctx.to(dtype))
    if with_vace:
        vc = torch.ones(b, 96, f, h, w)
        vc[:, :32] = torch.randn(b, 32, f, h, w, generator=g)
        out["vace_context"] = vc.to(dtype)
    if pin and torch.cuda.is_available():
        out = {k: v.pin_memory() for k, v in out.items()}
    return out


def model_flops(size: str, n_tokens: int, with_vace: bool, ctx_len: int = 512, num_layers: Optional[int] = None) -> float:
    """Algorithmic FLOPs of one model_fn call: per block
    8ND^2 + 4N^2D + 4ND^2 + 4*L*D^2 + 4*N*L*D + 4NDF; VACE adds its blocks + (n_vace + 1) D x D projections."""
    cfg = DIT_CONFIGS[size]
    d, f = cfg["dim"], cfg["ffn_dim"]
    layers = cfg["num_layers"] if num_layers is None else num_layers
    n, l = float(n_tokens), float(ctx_len)
    block = 8 * n * d * d + 4 * n * n * d + 4 * n * d * d + 4 * l * d * d + 4 * n * l * d + 4 * n * d * f
    total = layers * block
    if with_vace:
        nv = len([x for x in VACE_CONFIGS[size]["vace_layers"] if x < layers])
        total += nv * block + (nv + 1) * 2 * n * d * d + 2 * n * d * 384
    total += 2 * n * d * 64 * 2 + 2 * l * d * (cfg["text_dim"] + d)      # patch embed + head, text embedding
    return total


def attention_flops(n_tokens: int, heads: int, sk: Optional[int] = None) -> float:
    return 4.0 * n_tokens * (n_tokens if sk is None else sk) * heads * 128
