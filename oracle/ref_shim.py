"""Import the REAL reference modules from /root/reference (build container only)  --  TEST INFRASTRUCTURE.

Recipe of SURVEY.md Appendix A: mock the absent host-I/O dependencies, register stub packages whose
``__path__`` points at the reference directories, and import ``diffsynth.pipelines.wan_video_new``.
Nothing here is copied from the reference; nothing here can run on the GPU box (no /root/reference there).
"""
import importlib
import os
import sys
import types
from unittest.mock import MagicMock

REF_ROOT = os.environ.get("WVD_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "diffsynth"))


def load():
    """Returns (wan_video_new, wan_video_dit, wan_video_vace) real reference modules."""
    if "diffsynth.pipelines.wan_video_new" in sys.modules and hasattr(sys.modules["diffsynth.pipelines.wan_video_new"], "model_fn_wan_video"):
        w = sys.modules["diffsynth.pipelines.wan_video_new"]
        return w, sys.modules["diffsynth.models.wan_video_dit"], sys.modules["diffsynth.models.wan_video_vace"]
    for m in ("modelscope", "imageio", "ftfy"):
        try:
            importlib.import_module(m)
        except Exception:
            sys.modules[m] = MagicMock()
    root = os.path.join(REF_ROOT, "diffsynth")
    for name, sub in (("diffsynth", ""), ("diffsynth.models", "models"), ("diffsynth.pipelines", "pipelines"),
                      ("diffsynth.prompters", "prompters")):
        mod = types.ModuleType(name)
        mod.__path__ = [os.path.join(root, sub)]
        sys.modules[name] = mod
    sys.modules["diffsynth.prompters"].WanPrompter = MagicMock()
    sys.modules["diffsynth.models"].ModelManager = MagicMock()
    utils = importlib.import_module("diffsynth.models.utils")
    sys.modules["diffsynth.models"].load_state_dict = utils.load_state_dict
    w = importlib.import_module("diffsynth.pipelines.wan_video_new")
    dit = sys.modules["diffsynth.models.wan_video_dit"]
    vace = sys.modules["diffsynth.models.wan_video_vace"]
    dit.FLASH_ATTN_2_AVAILABLE = False   # FA2 is installed but cannot run on CPU/fp32 -> SDPA branch
    dit.FLASH_ATTN_3_AVAILABLE = False
    dit.SAGE_ATTN_AVAILABLE = False
    return w, dit, vace
