"""LoRA merge at load: ``W <- W + alpha * (B @ A)`` for every module the LoRA state dict names.

Host-side mirror of the reference's ``GeneralLoRALoader`` (diffsynth/lora/__init__.py:4-45; Ditto's rank-128 LoRA on
``vace_blocks.*.{self_attn,cross_attn}.{q,k,v,o}`` and ``ffn.{0,2}``, inference/infer_ditto.py:26): same class name,
same key parsing (``<module>.lora_B[.<adapter>].weight`` / ``.lora_A.``, optional ``diffusion_model.`` prefix), same
arithmetic (the product and the sum are evaluated in ``torch_dtype`` on ``device``), same 1x1-conv handling.  The
merge happens once at load, so the B200 kernels only ever see the merged weights; the only difference to the
reference is that the parameter is updated in place instead of round-tripping through ``load_state_dict``.
"""
from __future__ import annotations

import torch


class GeneralLoRALoader:
    def __init__(self, device="cpu", torch_dtype=torch.float32):
        self.device = device
        self.torch_dtype = torch_dtype

    def get_name_dict(self, lora_state_dict):
        names = {}
        for key in lora_state_dict:
            if ".lora_B." not in key:
                continue
            parts = key.split(".")
            i = parts.index("lora_B")
            if len(parts) > i + 2:
                parts.pop(i + 1)                 # adapter name ("default")
            parts.pop(parts.index("lora_B"))
            if parts[0] == "diffusion_model":
                parts.pop(0)
            parts.pop(-1)                        # "weight"
            names[".".join(parts)] = (key, key.replace(".lora_B.", ".lora_A."))
        return names

    @torch.no_grad()
    def load(self, model: torch.nn.Module, state_dict_lora, alpha=1.0) -> int:
        names = self.get_name_dict(state_dict_lora)
        updated = 0
        for name, module in model.named_modules():
            if name not in names:
                continue
            up = state_dict_lora[names[name][0]].to(device=self.device, dtype=self.torch_dtype)
            down = state_dict_lora[names[name][1]].to(device=self.device, dtype=self.torch_dtype)
            if up.dim() == 4:
                delta = alpha * torch.mm(up.squeeze(3).squeeze(2), down.squeeze(3).squeeze(2)).unsqueeze(2).unsqueeze(3)
            else:
                delta = alpha * torch.mm(up, down)
            w = module.weight
            merged = w.data.to(device=self.device, dtype=self.torch_dtype) + delta
            w.data.copy_(merged.to(device=w.device, dtype=w.dtype))
            updated += 1
        print(f"{updated} tensors are updated by LoRA.")
        return updated


def load_lora(module: torch.nn.Module, state_dict_lora, alpha: float = 1.0, device=None, torch_dtype=None) -> int:
    """``pipe.load_lora(module, path, alpha)`` with the state dict already read (wan_video_new.py / base pipeline):
    merges in the module's own dtype on its own device unless told otherwise."""
    p = next(module.parameters())
    return GeneralLoRALoader(device=device or p.device, torch_dtype=torch_dtype or p.dtype).load(module, state_dict_lora, alpha)
