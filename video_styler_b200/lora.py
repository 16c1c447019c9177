"""LoRA merge at load: ``W <- W + alpha * (B @ A)`` for every module the LoRA state dict names.

Host-side counterpart of the reference's ``GeneralLoRALoader`` (diffsynth/lora/__init__.py:4-45; Ditto's rank-128 LoRA
on ``vace_blocks.*.{self_attn,cross_attn}.{q,k,v,o}`` and ``ffn.{0,2}``, inference/infer_ditto.py:26).  Same public
surface -- ``GeneralLoRALoader(device, torch_dtype)``, ``get_name_dict``, ``load(model, state_dict_lora, alpha)`` -- and
the same arithmetic (product and sum evaluated in ``torch_dtype`` on ``device``, 1x1-conv factors squeezed to matrices),
checked bit for bit against the real loader in tests/test_install_reference.py.  The merge happens once at load, so the
B200 kernels only ever see merged weights.  Implementation differences: the targets are resolved through one
``dict(model.named_modules())`` lookup per LoRA pair, and the parameter is updated in place instead of round-tripping
through ``state_dict()`` / ``load_state_dict()``.
"""
from __future__ import annotations

from typing import Dict, Mapping, Tuple

import torch

_UP, _DOWN = ".lora_B.", ".lora_A."


def _target_of(up_key: str) -> str:
    """'[diffusion_model.]<module path>.lora_B[.<adapter>].weight' -> '<module path>' (reference :11-25)."""
    fields = up_key.split(".")
    at = fields.index("lora_B")
    path = fields[:at]                       # everything before 'lora_B'; adapter name and 'weight' come after it
    if path and path[0] == "diffusion_model":
        path = path[1:]
    return ".".join(path)


def _delta(up: torch.Tensor, down: torch.Tensor, alpha: float) -> torch.Tensor:
    """alpha * (B @ A), keeping the (out, in, 1, 1) shape of 1x1-conv LoRA factors (reference :35-40)."""
    if up.dim() == 4:
        return (alpha * torch.mm(up.flatten(1), down.flatten(1)))[:, :, None, None]
    return alpha * torch.mm(up, down)


class GeneralLoRALoader:
    def __init__(self, device="cpu", torch_dtype=torch.float32):
        self.device = device
        self.torch_dtype = torch_dtype

    def get_name_dict(self, lora_state_dict: Mapping[str, torch.Tensor]) -> Dict[str, Tuple[str, str]]:
        """module path -> (key of B / 'up', key of A / 'down')."""
        return {_target_of(k): (k, k.replace(_UP, _DOWN)) for k in lora_state_dict if _UP in k}

    @torch.no_grad()
    def load(self, model: torch.nn.Module, state_dict_lora: Mapping[str, torch.Tensor], alpha: float = 1.0) -> int:
        modules = dict(model.named_modules())
        cast = dict(device=self.device, dtype=self.torch_dtype)
        merged = 0
        for target, (up_key, down_key) in self.get_name_dict(state_dict_lora).items():
            module = modules.get(target)
            if module is None:               # the reference silently skips names the model does not have
                continue
            weight = module.weight
            new = weight.data.to(**cast) + _delta(state_dict_lora[up_key].to(**cast), state_dict_lora[down_key].to(**cast), alpha)
            weight.data.copy_(new.to(device=weight.device, dtype=weight.dtype))
            merged += 1
        print(f"{merged} tensors are updated by LoRA.")
        return merged


def load_lora(module: torch.nn.Module, state_dict_lora, alpha: float = 1.0, device=None, torch_dtype=None) -> int:
    """``pipe.load_lora(module, path, alpha)`` with the state dict already read: merges in the module's own dtype on
    its own device unless told otherwise."""
    p = next(module.parameters())
    return GeneralLoRALoader(device=device or p.device, torch_dtype=torch_dtype or p.dtype).load(module, state_dict_lora, alpha)
