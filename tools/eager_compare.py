"""Developer tool: the wvd path against a PyTorch-eager bf16 restatement of the same step ON THE SAME B200.

The reference itself cannot run on the GPU box (no /root/reference there), so the comparison partner is the oracle's
functional restatement executed on the GPU in bf16: F.linear (cuBLAS), F.scaled_dot_product_attention (cuDNN / flash),
torch LayerNorm / RMSNorm / RoPE as separate eager kernels -- i.e. what the reference's own code path launches, minus
its Python module overhead.  c3 shapes at reduced depth (the oracle keeps every intermediate alive); reports seconds
per call and the parity of the two outputs.  Not part of bench.py: the oracle is a checker, never the product.

    python tools/eager_compare.py [--layers 6]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import video_styler_b200 as V  # noqa: E402
from oracle import wan_oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=6)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
DEV = "cuda"
cfg = dict(O.DIT_CONFIGS["14B"]); cfg["num_layers"] = a.layers
vcfg = dict(O.VACE_CONFIGS["14B"]); vcfg["vace_layers"] = tuple(l for l in vcfg["vace_layers"] if l < a.layers)
sd = O.make_state_dict(O.dit_param_shapes(cfg), seed=0, device=DEV, dtype=torch.bfloat16)
vsd = O.make_state_dict(O.vace_param_shapes(vcfg), seed=3, device=DEV, dtype=torch.bfloat16)
O.lora_merge(vsd, O.make_lora_state_dict(vcfg, seed=2, rank=128, device=DEV, dtype=torch.bfloat16))
with torch.device("meta"):
    dit = V.WanModel(has_image_input=False, **cfg)
    vace = V.VaceWanModel(has_image_input=False, **vcfg)
dit.load_state_dict(sd, strict=True, assign=True)
vace.load_state_dict(vsd, strict=True, assign=True)
dit.freqs = V.wan_video_dit.precompute_freqs_cis_3d(128)
dit.requires_grad_(False), vace.requires_grad_(False)
inp = O.make_inputs((1, 16, 19, 60, 104), cfg["text_dim"], seed=1, with_vace=True, device=DEV, dtype=torch.bfloat16)
ts = torch.tensor([832.0], device=DEV, dtype=torch.bfloat16)


def ours():
    return V.model_fn_wan_video(dit=dit, vace=vace, latents=inp["latents"], timestep=ts, context=inp["context"],
                                vace_context=inp["vace_context"], vace_scale=1.0)


def eager():
    return O.model_fn_wan_video(sd, cfg, inp["latents"], ts, inp["context"], vsd, vcfg, inp["vace_context"], 1.0)


def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.reps / 1e3, out


with torch.no_grad():
    t_o, out_o = timed(ours)
    t_e, out_e = timed(eager)
blocks = a.layers + len(vcfg["vace_layers"])
print(f"c3 shapes, {a.layers} main + {len(vcfg['vace_layers'])} VACE blocks ({blocks} DiT blocks): wvd {t_o:.4f} s  torch-eager bf16 {t_e:.4f} s  "
      f"ratio {t_e / t_o:.2f}x   per block: wvd {t_o / blocks * 1e3:.1f} ms, eager {t_e / blocks * 1e3:.1f} ms")
print("parity wvd vs eager:", O.parity_metrics(out_o, out_e))
