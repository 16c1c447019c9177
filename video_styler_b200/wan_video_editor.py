"""Keyframe-guided flow-matching video editing on the B200 path -- the fork's namesake feature.

Mirrors ``diffsynth/pipelines/wan_video_editor.py`` (WanVideoEditorPipeline): same method names, argument meaning and
error behaviour for the arithmetic between the VAE / text encoder and the DiT --

    prepare_coupled_noise        :47-75     the edited keyframes start from the SAME noise as their source frames
    construct_rope_ids           :77-105    temporal RoPE ids [0 .. T-1 | keyframe_indices]: an edited keyframe shares
                                            the position encoding of the frame it replaces
    compute_velocity_correction  :107-165   dv at the keyframe positions from the consistency residual r_k
    compute_metrics              :167-196   monitoring numbers
    the denoising loop           :352-392   joint DiT call on cat([z_main, z_edit]) with ``rope_indices``, CFG, split,
                                            correction, Euler step of both latent sets

-- with the DiT call on ``model_fn_wan_video(rope_indices=...)`` (the reference's own ``WanModel.forward`` is broken
in the snapshot, SURVEY.md 0.2: ``wan_video_dit.py:375`` unpacks a tuple ``patchify`` does not return) and everything
after it fused into ONE kernel per step (``wvd_editor_step``: CFG combine + correction + Euler for both latent sets,
reference rounding points), instead of ~20 elementwise launches and two ``.item()`` synchronisations per step.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import engine, ops as _cuda_ops
from .pipeline import FlowMatchScheduler, model_fn_wan_video

Tensor = torch.Tensor


def generate_noise(shape, seed=None, device="cpu", dtype=torch.float16) -> Tensor:
    """BasePipeline.generate_noise (diffsynth/pipelines/base.py:124-127)."""
    generator = None if seed is None else torch.Generator(device).manual_seed(seed)
    return torch.randn(shape, generator=generator, device=device, dtype=dtype)


def prepare_coupled_noise(latent_shape: Tuple[int, ...], keyframe_indices: Sequence[int], seed: Optional[int] = None,
                          device: str = "cpu") -> Tuple[Tensor, Tensor]:
    """fp32 noise for the main latents and, cut out of it, the noise of the edited keyframes (wan_video_editor.py:47-75)."""
    noise_main = generate_noise(latent_shape, seed=seed, device=device, dtype=torch.float32)
    noise_edit = noise_main[:, :, list(keyframe_indices), :, :].clone()
    return noise_main, noise_edit


def construct_rope_ids(total_frames: int, keyframe_indices: Sequence[int], device="cuda") -> Tensor:
    """[0 .. T-1 | keyframe_indices] (wan_video_editor.py:77-105)."""
    ids_main = torch.arange(total_frames, device=device)
    ids_edit = torch.tensor(list(keyframe_indices), device=device)
    return torch.cat([ids_main, ids_edit])


class KeyframeMap:
    """keyframe_indices as the two device lookup tables the kernel reads: frame -> keyframe slot (or -1), slot -> frame."""

    def __init__(self, total_frames: int, keyframe_indices: Sequence[int], device):
        idx = [int(i) for i in keyframe_indices]
        if len(idx) == 0:
            raise ValueError("keyframe_indices is empty")
        if len(set(idx)) != len(idx):
            # the reference's `v[:, :, keyframe_indices] += correction` keeps an unspecified one of the duplicates
            raise ValueError(f"keyframe_indices must be unique, got {idx}")
        if min(idx) < 0 or max(idx) >= total_frames:
            raise IndexError(f"keyframe index out of range for {total_frames} latent frames: {idx}")
        f2k = [-1] * total_frames
        for k, t in enumerate(idx):
            f2k[t] = k
        self.indices = idx
        self.total_frames = total_frames
        self.frame_to_key = torch.tensor(f2k, dtype=torch.int32, device=device)
        self.key_idx = torch.tensor(idx, dtype=torch.int32, device=device)


def compute_velocity_correction(z_main: Tensor, z_edit: Tensor, v_main: Tensor, v_edit: Tensor,
                                keyframe_indices: Sequence[int], dt: float, alpha: float = 10.0, beta: float = 0.0,
                                ops=_cuda_ops) -> Tuple[Tensor, Tensor]:
    """(v_main_corrected, v_edit_corrected) -- wan_video_editor.py:107-165, one kernel."""
    km = KeyframeMap(z_main.shape[2], keyframe_indices, z_main.device)
    return ops.editor_step(z_main.contiguous(), z_edit.contiguous(), (v_main.contiguous(), v_edit.contiguous()), None,
                           km.frame_to_key, km.key_idx, 1.0, dt, alpha, beta, 0.0, euler=False)


def compute_metrics(z_main: Tensor, z_edit: Tensor, v_main: Tensor, v_edit: Tensor, keyframe_indices: Sequence[int],
                    dt: float) -> Dict[str, float]:
    """Monitoring numbers of wan_video_editor.py:167-196 (synchronises: the reference prints them every 10th step)."""
    idx = list(keyframe_indices)
    z_diff = z_main[:, :, idx] - z_edit
    v_diff = v_main[:, :, idx] - v_edit
    r_k = z_diff - v_diff * dt
    return {"r_k_norm": torch.mean(torch.abs(r_k)).item(), "v_diff_norm": torch.mean(torch.abs(v_diff)).item(),
            "delta_v_norm": torch.mean(torch.abs(z_diff)).item()}


@torch.no_grad()
def edit_denoise(dit, z_main: Tensor, z_edit: Tensor, context_posi: Tensor, context_nega: Optional[Tensor],
                 keyframe_indices: Sequence[int], num_inference_steps: int = 50, cfg_scale: float = 5.0,
                 sigma_shift: float = 5.0, alpha: float = 10.0, beta: float = 0.0, torch_dtype=torch.bfloat16,
                 scheduler: Optional[FlowMatchScheduler] = None, progress=None, verbose: bool = False,
                 cache_text: bool = True, ops=_cuda_ops) -> Tuple[Tensor, Tensor]:
    """The denoising loop of WanVideoEditorPipeline.__call__ (wan_video_editor.py:352-392) from the coupled noise
    (z_main, z_edit) to the final latents; VAE encode / decode and prompt encoding stay with the caller.

    Per step: ONE joint DiT call per CFG branch on cat([z_main, z_edit], dim=2) with rope ids
    [0..T-1 | keyframe_indices], then ``ops.editor_step`` (CFG + split + velocity correction + Euler, one kernel).
    dt follows the reference: timesteps[i] - timesteps[i+1] in fp32, 0 at the last step; the Euler factor is the
    scheduler's sigma difference."""
    sch = scheduler or FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True)
    sch.set_timesteps(num_inference_steps, denoising_strength=1.0, shift=sigma_shift)
    km = KeyframeMap(z_main.shape[2], keyframe_indices, z_main.device)
    rope_ids = construct_rope_ids(z_main.shape[2], km.indices, device=z_main.device)
    text_cache = engine.TextCache() if cache_text else None
    steps = list(range(len(sch.timesteps)))
    for i in (steps if progress is None else progress(steps)):
        ts = sch.timesteps[i]
        timestep = ts.unsqueeze(0).to(dtype=torch_dtype, device=z_main.device)
        z_concat = torch.cat([z_main, z_edit], dim=2)
        kw = dict(dit=dit, latents=z_concat, timestep=timestep, rope_indices=rope_ids, ops=ops)
        if text_cache is not None:
            kw["text_cache"] = text_cache
        v_posi = model_fn_wan_video(context=context_posi, **kw)
        v_nega = model_fn_wan_video(context=context_nega, **kw) if cfg_scale != 1.0 else None
        dt = float(sch.timesteps[i] - sch.timesteps[i + 1]) if i < len(sch.timesteps) - 1 else 0.0
        if verbose and i % 10 == 0:
            v = v_posi if v_nega is None else v_nega + cfg_scale * (v_posi - v_nega)
            m = compute_metrics(z_main, z_edit, v[:, :, :z_main.shape[2]], v[:, :, z_main.shape[2]:], km.indices, dt)
            print(f"Step {i}: r_k={m['r_k_norm']:.6f}, v_diff={m['v_diff_norm']:.6f}, dv={m['delta_v_norm']:.6f}")
        z_main, z_edit = ops.editor_step(z_main.contiguous(), z_edit.contiguous(), v_posi.contiguous(),
                                         None if v_nega is None else v_nega.contiguous(), km.frame_to_key, km.key_idx,
                                         cfg_scale, dt, alpha, beta, sch.dsigma(ts), euler=True)
    return z_main, z_edit
