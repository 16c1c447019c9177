"""Drop-in check against the REAL reference objects (runs only where /root/reference exists, i.e. in the build
container): install() on a reference-style pipeline holding the real WanModel / VaceWanModel -- with the real
GeneralLoRALoader merge and the real enable_vram_management wrappers applied -- must reproduce the reference output.
The kernels are replaced by the CPU stand-in so this exercises exactly the host-side boundary logic."""
import types

import pytest
import torch

from oracle import ref_shim, wan_oracle as O
from tests import cpu_backend

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="needs /root/reference (build container only)")


def test_install_on_real_reference_modules_with_vram_wrappers(golden_dir):
    import functools
    import os
    import video_styler_b200 as V
    w, dit_mod, vace_mod = ref_shim.load()
    from diffsynth.lora import GeneralLoRALoader
    from diffsynth.vram_management import AutoWrappedLinear, AutoWrappedModule, WanAutoCastLayerNorm, enable_vram_management
    fix = torch.load(os.path.join(golden_dir, "tiny_vace_lora.pt"), weights_only=False)
    cfg, vcfg = O.DIT_CONFIGS["tiny"], O.VACE_CONFIGS["tiny"]
    dit = dit_mod.WanModel(has_image_input=False, **cfg)
    dit.load_state_dict(O.make_state_dict(O.dit_param_shapes(cfg), seed=0, perturb_norms=True), strict=True)
    vace = vace_mod.VaceWanModel(has_image_input=False, **vcfg)
    vace.load_state_dict(O.make_state_dict(O.vace_param_shapes(vcfg), seed=3, perturb_norms=True), strict=True)
    GeneralLoRALoader(device="cpu", torch_dtype=torch.float32).load(vace, O.make_lora_state_dict(vcfg, seed=2, rank=16), alpha=1.0)
    mcfg = dict(offload_dtype=torch.float32, offload_device="cpu", onload_dtype=torch.float32, onload_device="cpu",
                computation_dtype=torch.float32, computation_device="cpu")
    # the module maps of WanVideoPipeline.enable_vram_management (wan_video_new.py:152-184, 272-291)
    enable_vram_management(dit, module_map={torch.nn.Linear: AutoWrappedLinear, torch.nn.Conv3d: AutoWrappedModule,
                                            torch.nn.LayerNorm: WanAutoCastLayerNorm, dit_mod.RMSNorm: AutoWrappedModule},
                           module_config=mcfg)
    enable_vram_management(vace, module_map={torch.nn.Linear: AutoWrappedLinear, torch.nn.Conv3d: AutoWrappedModule,
                                             torch.nn.LayerNorm: AutoWrappedModule, dit_mod.RMSNorm: AutoWrappedModule},
                           module_config=mcfg)
    assert type(dit.blocks[0].self_attn.q).__name__ == "AutoWrappedLinear"
    assert type(vace.vace_blocks[0].norm1).__name__ == "AutoWrappedModule"
    pipe = types.SimpleNamespace(model_fn=w.model_fn_wan_video, dit=dit, vace=vace)
    V.install(pipe)
    assert pipe.model_fn.func is V.model_fn_wan_video and isinstance(pipe.wvd_text_cache, V.TextCache)
    inp = O.make_inputs(fix["latent_shape"], cfg["text_dim"], seed=1, with_vace=True)
    fn = functools.partial(pipe.model_fn, ops=cpu_backend)
    with torch.no_grad():
        out = fn(dit=pipe.dit, vace=pipe.vace, latents=inp["latents"], timestep=torch.tensor([fix["timestep"]]),
                 context=inp["context"], vace_context=inp["vace_context"], vace_scale=1.0,
                 tea_cache=None, use_unified_sequence_parallel=False, motion_bucket_id=None, cfg_merge=False)
    m = O.parity_metrics(out, fix["output"])
    assert m["max_abs"] <= 5e-5 and m["rel_l2"] <= 2e-5, m
    # second call with the same context object: text embedding + cross K/V come from the cache, same bits
    with torch.no_grad():
        out2 = fn(dit=pipe.dit, vace=pipe.vace, latents=inp["latents"], timestep=torch.tensor([fix["timestep"]]),
                  context=inp["context"], vace_context=inp["vace_context"], vace_scale=1.0)
    assert torch.equal(out, out2) and pipe.wvd_text_cache.hits == 1 and pipe.wvd_text_cache.misses == 1


def test_lora_loader_equals_the_real_general_lora_loader():
    """video_styler_b200.GeneralLoRALoader against the reference's own class on the reference's own VaceWanModel:
    identical merged weights, bit for bit (fp32 and bf16 merge dtypes)."""
    import video_styler_b200 as V
    w, dit_mod, vace_mod = ref_shim.load()
    from diffsynth.lora import GeneralLoRALoader
    vcfg = O.VACE_CONFIGS["tiny"]
    lsd = O.make_lora_state_dict(vcfg, seed=2, rank=16)
    for dt in (torch.float32, torch.bfloat16):
        sd = O.make_state_dict(O.vace_param_shapes(vcfg), seed=3, perturb_norms=True)
        ref = vace_mod.VaceWanModel(has_image_input=False, **vcfg)
        ref.load_state_dict(sd, strict=True)
        ours = V.VaceWanModel(has_image_input=False, **vcfg)
        ours.load_state_dict(sd, strict=True)
        GeneralLoRALoader(device="cpu", torch_dtype=dt).load(ref, lsd, alpha=0.5)
        V.GeneralLoRALoader(device="cpu", torch_dtype=dt).load(ours, lsd, alpha=0.5)
        a, b = ref.state_dict(), ours.state_dict()
        assert a.keys() == b.keys()
        for k in a:
            assert torch.equal(a[k], b[k]), (k, dt)


def test_install_vae_on_the_real_reference_vae_object():
    """install_vae() on a REAL WanVideoVAE object (its convolutional model replaced by the stand-in so that the test runs
    in seconds): the rebound encode / decode reproduce the real methods bit for bit, fp32 and bf16."""
    import importlib
    from oracle import aux_oracle as A
    from video_styler_b200 import wan_video_vae as VA
    ref_shim.load()
    R = importlib.import_module("diffsynth.models.wan_video_vae")

    def real_vae():
        vae = object.__new__(R.WanVideoVAE)
        torch.nn.Module.__init__(vae)
        vae.model, vae.upsampling_factor, vae.z_dim = A.ToyVAEModel(), 8, 16
        vae.mean, vae.std = torch.zeros(16), torch.ones(16)
        vae.scale = [vae.mean, 1.0 / vae.std]
        return vae
    for dt in (torch.float32, torch.bfloat16):
        z = A.make_vae_source((16, 3, 9, 13), dtype=dt)
        x = A.make_vae_source((3, 9, 72, 104), dtype=dt)
        ref = real_vae()
        with torch.no_grad():
            want_dec = ref.decode([z], "cpu", tiled=True, tile_size=(4, 6), tile_stride=(2, 3))
            want_enc = ref.encode([x], "cpu", tiled=True, tile_size=(4, 6), tile_stride=(2, 3))
            want_single = ref.decode([z], "cpu", tiled=False)
        pipe = types.SimpleNamespace(vae=real_vae())
        VA.install_vae(pipe, ops=cpu_backend)
        with torch.no_grad():
            assert torch.equal(pipe.vae.decode([z], "cpu", tiled=True, tile_size=(4, 6), tile_stride=(2, 3)), want_dec)
            assert torch.equal(pipe.vae.encode([x], "cpu", tiled=True, tile_size=(4, 6), tile_stride=(2, 3)), want_enc)
            assert torch.equal(pipe.vae.decode([z], "cpu", tiled=False), want_single)


def test_text_encoder_loads_the_real_reference_state_dict_and_matches_it():
    """The REAL WanTextEncoder's own state_dict() (its init_weights, its key names) loads with strict=True; forward through
    the CPU stand-in of the kernels reproduces the real module's output."""
    import importlib
    from oracle import aux_oracle as A
    from video_styler_b200 import wan_video_text_encoder as T
    ref_shim.load()
    R = importlib.import_module("diffsynth.models.wan_video_text_encoder")
    cfg = A.T5_CONFIGS["tiny"]
    torch.manual_seed(0)
    ref = R.WanTextEncoder(**cfg).eval()
    for blk in ref.blocks:                       # the reference's init makes q tiny: give the softmax something to do
        blk.attn.q.weight.data.mul_(40.0)
    ours = T.WanTextEncoder(**cfg).eval().requires_grad_(False)
    res = ours.load_state_dict(ref.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    ids, mask = A.make_t5_inputs(cfg, 33, 20)
    with torch.no_grad():
        m = O.parity_metrics(ours(ids, mask, ops=cpu_backend), ref(ids, mask))
    assert m["rel_l2"] <= 1e-5, m
