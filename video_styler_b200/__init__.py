"""video_styler_b200 -- B200-native (sm_100a) Wan2.1(-VACE) DiT denoising forward for Ditto / Editto.

Host code is Python/PyTorch (device memory, streams, torch.distributed); the math runs in libwvd.so, a C-ABI
library of hand-written CUDA kernels (include/wvd.h).  Public surface mirrors the reference:

    model_fn_wan_video, WanModel, VaceWanModel, FlowMatchScheduler, GeneralLoRALoader / load_lora, denoise, install(pipe)

and, for the callers either side of the DiT: wan_video_text_encoder.WanTextEncoder (umT5),
wan_video_editor (keyframe editing loop), wan_video_vae.TiledVAE / install_vae (VAE tiling layer).
"""
from ._lib import WvdError  # noqa: F401
from .lora import GeneralLoRALoader, load_lora  # noqa: F401
from .engine import TextCache  # noqa: F401
from .pipeline import FlowMatchScheduler, GraphedModelFn, denoise, install, model_fn_wan_video  # noqa: F401
from .wan_video_dit import WanModel  # noqa: F401
from .wan_video_vace import VaceWanModel  # noqa: F401
from .wan_video_text_encoder import WanTextEncoder  # noqa: F401
from .wan_video_vae import TiledVAE, install_vae  # noqa: F401
from .wan_video_editor import edit_denoise  # noqa: F401

__version__ = "0.1.0"
