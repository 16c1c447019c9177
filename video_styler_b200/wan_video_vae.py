"""The tiling layer of the Wan VAE on the B200 path (PARTIAL: see below).

Mirrors the methods of ``WanVideoVAE`` that the pipeline calls -- ``encode`` / ``decode`` / ``tiled_encode`` /
``tiled_decode`` / ``single_encode`` / ``single_decode`` / ``build_1d_mask`` / ``build_mask``
(diffsynth/models/wan_video_vae.py:1081-1248): same tile enumeration, the same linear-ramp masks, the same
accumulate-then-normalise blending with the reference's rounding -- with two differences in HOW:

  * the reference keeps ``values`` / ``weight`` on the CPU and ships every tile's input and output over PCIe
    (``data_device = "cpu"``, :1118, :1170); here the video, the latents and both accumulators stay in HBM;
  * per tile the reference materialises the mask, multiplies, adds into two strided windows (6 tensor ops, 2 of them over
    the full (1, C, T, h, w) tile); here one kernel (``wvd_tile_blend``) computes the mask analytically and updates both
    windows, and one kernel (``wvd_tile_finalize``) divides and clamps.

What is NOT re-implemented: the convolutional model behind ``self.model.encode / decode`` (``VideoVAE_``: causal 3-D
convolutions with a chunked feature cache, :951-1056) -- it is taken as given (the reference's own torch module on the same
GPU) and called per tile exactly like the reference does.  DESIGN.md section 7 has the plan for it.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from . import ops as _cuda_ops

Tensor = torch.Tensor

# latent statistics of the Wan2.1 VAE (wan_video_vae.py:1063-1072): part of the checkpoint contract, not code
LATENT_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632,
               -0.1922, -0.9497, 0.2503, -0.2921]
LATENT_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382,
              1.1253, 2.8251, 1.9160]


def tile_tasks(height: int, width: int, tile_size: Tuple[int, int], tile_stride: Tuple[int, int]) -> List[Tuple[int, int, int, int]]:
    """(h, h_end, w, w_end) windows in the reference's order (:1108-1116): a window is dropped when the one before it
    already reaches the far edge."""
    (size_h, size_w), (stride_h, stride_w) = tile_size, tile_stride
    tasks = []
    for h in range(0, height, stride_h):
        if h - stride_h >= 0 and h - stride_h + size_h >= height:
            continue
        for w in range(0, width, stride_w):
            if w - stride_w >= 0 and w - stride_w + size_w >= width:
                continue
            tasks.append((h, h + size_h, w, w + size_w))
    return tasks


class TiledVAE:
    """``WanVideoVAE``'s encode / decode surface around a given convolutional model (``model.encode(x, scale)``,
    ``model.decode(z, scale)``), blending on the device.  ``install_vae(pipe)`` rebinds a reference pipeline's VAE."""

    def __init__(self, model, z_dim: int = 16, upsampling_factor: int = 8, mean: Sequence[float] = LATENT_MEAN,
                 std: Sequence[float] = LATENT_STD, ops=_cuda_ops):
        self.model, self.z_dim, self.upsampling_factor, self.ops = model, z_dim, upsampling_factor, ops
        self.mean, self.std = torch.tensor(list(mean)), torch.tensor(list(std))
        self.scale = [self.mean, 1.0 / self.std]

    # -- the reference's mask helpers, for callers that use them (:1081-1100) --
    def build_1d_mask(self, length, left_bound, right_bound, border_width):
        x = torch.ones((length,))
        ramp = (torch.arange(border_width) + 1) / max(border_width, 1)
        if not left_bound:
            x[:border_width] = ramp
        if not right_bound:
            x[-border_width:] = torch.flip(ramp, dims=(0,))
        return x

    def build_mask(self, data, is_bound, border_width):
        hh, ww = data.shape[3], data.shape[4]
        mh = self.build_1d_mask(hh, is_bound[0], is_bound[1], border_width[0])
        mw = self.build_1d_mask(ww, is_bound[2], is_bound[3], border_width[1])
        return torch.minimum(mh[:, None].expand(hh, ww), mw[None, :].expand(hh, ww)).reshape(1, 1, 1, hh, ww)

    def _blend(self, source: Tensor, device, tile_size, tile_stride, run, out_channels: int, out_t: int, to_out, border, clamp):
        """Shared body of tiled_decode / tiled_encode: enumerate the windows of ``source``'s (H, W), run the model on each,
        blend into (1, out_channels, out_t, to_out(H), to_out(W))."""
        _, _, _, hgt, wid = source.shape
        source = source.to(device)
        out_h, out_w = to_out(hgt), to_out(wid)
        values = torch.zeros((1, out_channels, out_t, out_h, out_w), dtype=source.dtype, device=device)
        weight = torch.zeros((out_h, out_w), dtype=source.dtype, device=device)
        for h, h_, w, w_ in tile_tasks(hgt, wid, tile_size, tile_stride):
            tile = run(source[:, :, :, h:h_, w:w_]).to(dtype=source.dtype).contiguous()
            self.ops.tile_blend(values, weight, tile, to_out(h), to_out(w), (h == 0, h_ >= hgt, w == 0, w_ >= wid), border)
        return self.ops.tile_finalize(values, weight, clamp)

    def tiled_decode(self, hidden_states: Tensor, device, tile_size, tile_stride) -> Tensor:
        """:1103-1153.  Returns the video on ``device`` (the reference returns it on the CPU)."""
        f = self.upsampling_factor
        (size_h, size_w), (stride_h, stride_w) = tile_size, tile_stride
        return self._blend(hidden_states, device, tile_size, tile_stride, lambda z: self.model.decode(z, self.scale), 3,
                           hidden_states.shape[2] * 4 - 3, lambda v: v * f, ((size_h - stride_h) * f, (size_w - stride_w) * f),
                           (-1.0, 1.0))

    def tiled_encode(self, video: Tensor, device, tile_size, tile_stride) -> Tensor:
        """:1155-1204 (tile_size / tile_stride in pixels, as ``encode`` passes them)."""
        f = self.upsampling_factor
        (size_h, size_w), (stride_h, stride_w) = tile_size, tile_stride
        return self._blend(video, device, tile_size, tile_stride, lambda x: self.model.encode(x, self.scale), self.z_dim,
                           (video.shape[2] + 3) // 4, lambda v: v // f, ((size_h - stride_h) // f, (size_w - stride_w) // f), None)

    def single_encode(self, video: Tensor, device) -> Tensor:
        return self.model.encode(video.to(device), self.scale)

    def single_decode(self, hidden_state: Tensor, device) -> Tensor:
        return self.model.decode(hidden_state.to(device), self.scale).clamp_(-1, 1)

    def encode(self, videos, device, tiled=False, tile_size=(34, 34), tile_stride=(18, 16)) -> Tensor:
        """:1218-1232 -- ``tile_size`` / ``tile_stride`` in latent pixels."""
        f = self.upsampling_factor
        out = []
        for video in videos:
            video = video.unsqueeze(0)
            if tiled:
                hs = self.tiled_encode(video, device, (tile_size[0] * f, tile_size[1] * f), (tile_stride[0] * f, tile_stride[1] * f))
            else:
                hs = self.single_encode(video, device)
            out.append(hs.squeeze(0))
        return torch.stack(out)

    def decode(self, hidden_states, device, tiled=False, tile_size=(34, 34), tile_stride=(18, 16)) -> Tensor:
        """:1235-1248."""
        out = []
        for hs in hidden_states:
            hs = hs.unsqueeze(0)
            out.append((self.tiled_decode(hs, device, tile_size, tile_stride) if tiled else self.single_decode(hs, device)).squeeze(0))
        return torch.stack(out)


def install_vae(pipe, ops=_cuda_ops):
    """Rebind ``pipe.vae``'s encode / decode / tiled_* methods (a reference ``WanVideoVAE``) to the device-resident
    blending above; its convolutional ``model``, statistics and state dict stay untouched."""
    vae = pipe.vae
    t = TiledVAE(vae.model, z_dim=vae.z_dim, upsampling_factor=vae.upsampling_factor, mean=vae.mean.tolist(), std=vae.std.tolist(), ops=ops)
    for name in ("encode", "decode", "tiled_encode", "tiled_decode", "single_encode", "single_decode"):
        setattr(vae, name, getattr(t, name))
    vae.wvd_tiled = t
    return pipe
