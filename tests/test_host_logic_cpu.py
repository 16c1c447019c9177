"""Host-side logic without a GPU: the package's module tree (state-dict key parity with the reference), the
engine/pipeline orchestration (driven through a CPU stand-in for the kernel wrappers) against the golden vectors
from the REAL reference, the scheduler, and the loud failure when the CUDA path is unavailable."""
import os

import pytest
import torch

import video_styler_b200 as V
from oracle import wan_oracle as O
from tests import cpu_backend
from video_styler_b200 import _lib, engine, ops


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)


def build_models(fix, dtype=torch.float32, device="cpu"):
    cfg = O.DIT_CONFIGS[fix["size"]]
    dit = V.WanModel(has_image_input=False, **cfg)
    sd = O.make_state_dict(O.dit_param_shapes(cfg), seed=fix["seeds"]["dit"], perturb_norms=fix["perturb"],
                           weight_scale=fix["weight_scale"])
    res = dit.load_state_dict(sd, strict=True)           # key names == the reference's state-dict keys
    assert not res.missing_keys and not res.unexpected_keys
    vace = None
    if fix["with_vace"]:
        vcfg = O.VACE_CONFIGS[fix["size"]]
        vace = V.VaceWanModel(has_image_input=False, **vcfg)
        vsd = O.make_state_dict(O.vace_param_shapes(vcfg), seed=fix["seeds"]["vace"], perturb_norms=fix["perturb"],
                                weight_scale=fix["weight_scale"])
        if fix["lora"]:
            O.lora_merge(vsd, O.make_lora_state_dict(vcfg, seed=fix["seeds"]["lora"], rank=fix["lora_rank"]))
        vace.load_state_dict(vsd, strict=True)
        vace = vace.to(device=device, dtype=dtype).eval().requires_grad_(False)
    return dit.to(device=device, dtype=dtype).eval().requires_grad_(False), vace


def run_model_fn(fix, dit, vace, backend, device="cpu", dtype=torch.float32, **kw):
    cfg = O.DIT_CONFIGS[fix["size"]]
    inp = O.make_inputs(fix["latent_shape"], cfg["text_dim"], seed=fix["seeds"]["inputs"], with_vace=fix["with_vace"])
    ts = torch.tensor([fix["timestep"]], dtype=torch.float32).to(device=device, dtype=dtype)
    with torch.no_grad():
        return V.model_fn_wan_video(dit=dit, vace=vace, latents=inp["latents"].to(device=device, dtype=dtype),
                                    timestep=ts, context=inp["context"].to(device=device, dtype=dtype),
                                    vace_context=inp["vace_context"].to(device=device, dtype=dtype) if fix["with_vace"] else None,
                                    vace_scale=1.0, ops=backend, **kw)


@pytest.mark.parametrize("name", ["tiny_t2v", "tiny_vace_lora", "small_vace"])
def test_pipeline_orchestration_matches_reference_golden(golden_dir, name):
    fix = _load(golden_dir, name)
    dit, vace = build_models(fix)
    out = run_model_fn(fix, dit, vace, cpu_backend)
    m = O.parity_metrics(out, fix["output"])
    assert m["max_abs"] <= 5e-5 and m["rel_l2"] <= 2e-5, m


def test_state_dict_keys_match_reference_names():
    cfg, vcfg = O.DIT_CONFIGS["tiny"], O.VACE_CONFIGS["tiny"]
    assert set(V.WanModel(has_image_input=False, **cfg).state_dict()) == set(O.dit_param_shapes(cfg))
    assert set(V.VaceWanModel(**vcfg).state_dict()) == set(O.vace_param_shapes(vcfg))
    # GeneralLoRALoader walks named_modules() for 'vace_blocks.N.self_attn.q' ... (diffsynth/lora/__init__.py:31-42)
    names = dict(V.VaceWanModel(**vcfg).named_modules())
    for tgt in O.LORA_TARGETS:
        assert isinstance(names[f"vace_blocks.1.{tgt}"], torch.nn.Linear)


def test_state_dict_converters_infer_config():
    cfg, vcfg = O.DIT_CONFIGS["1.3B"], O.VACE_CONFIGS["14B"]
    shapes = O.dit_param_shapes(cfg)
    meta = {k: torch.empty(v, device="meta") for k, v in shapes.items()}
    meta["vace_patch_embedding.weight"] = torch.empty(1, device="meta")
    sd, got = V.WanModel.state_dict_converter().from_civitai({"model.diffusion_model." + k if not k.startswith("vace") else k: v
                                                              for k, v in meta.items()})
    assert set(sd) == set(shapes)
    for k in ("dim", "in_dim", "ffn_dim", "out_dim", "text_dim", "freq_dim", "num_heads", "num_layers", "patch_size"):
        assert got[k] == cfg[k], k
    vmeta = {k: torch.empty(v, device="meta") for k, v in O.vace_param_shapes(vcfg).items()}
    vmeta["blocks.0.modulation"] = torch.empty(1, device="meta")
    vsd, vgot = V.VaceWanModel.state_dict_converter().from_civitai(vmeta)
    assert set(vsd) == set(O.vace_param_shapes(vcfg))
    for k in ("vace_layers", "vace_in_dim", "dim", "num_heads", "ffn_dim"):
        assert vgot[k] == vcfg[k], k


def test_rope_indices_variant_matches_oracle():
    """The fixed WanModel.forward(rope_indices=...) of wan_video_dit.py:378-384."""
    fix = dict(size="tiny", seeds=dict(dit=0, vace=3, lora=2, inputs=1), perturb=True, weight_scale=1.0,
               with_vace=False, lora=False, latent_shape=(1, 16, 3, 8, 12), timestep=700.0)
    dit, _ = build_models(fix)
    idx = torch.tensor([0, 7, 20])
    out = run_model_fn(fix, dit, None, cpu_backend, rope_indices=idx)
    cfg = O.DIT_CONFIGS["tiny"]
    sd = O.make_state_dict(O.dit_param_shapes(cfg), seed=0, perturb_norms=True)
    inp = O.make_inputs(fix["latent_shape"], cfg["text_dim"], seed=1)
    with torch.no_grad():
        ref = O.model_fn_wan_video(sd, cfg, inp["latents"], torch.tensor([700.0]), inp["context"], rope_indices=idx)
    assert O.parity_metrics(out, ref)["max_abs"] <= 5e-5


def test_scheduler_and_denoise_loop(golden_dir):
    fix = _load(golden_dir, "flow_match")
    sch = V.FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True)
    sch.set_timesteps(50, shift=5.0)
    assert torch.equal(sch.sigmas, fix["sigmas"]) and torch.equal(sch.timesteps, fix["timesteps"])
    for i, ref in fix["steps"].items():
        assert torch.allclose(sch.step(fix["v"], sch.timesteps[i], fix["x"]), ref, atol=1e-6, rtol=0)
    # 3-step CFG loop through the package vs the oracle's restatement of wan_video_new.py:515-542
    f = dict(size="tiny", seeds=dict(dit=0, vace=3, lora=2, inputs=1), perturb=True, weight_scale=1.0,
             with_vace=False, lora=False)
    dit, _ = build_models(f)
    cfg = O.DIT_CONFIGS["tiny"]
    sd = O.make_state_dict(O.dit_param_shapes(cfg), seed=0, perturb_norms=True)
    inp = O.make_inputs((1, 16, 3, 8, 12), cfg["text_dim"], seed=1)
    nega = torch.zeros_like(inp["context"])
    import functools
    import video_styler_b200.pipeline as P
    orig = P.model_fn_wan_video
    P.model_fn_wan_video = functools.partial(orig, ops=cpu_backend)
    try:
        got = V.denoise(dit, None, inp["latents"], inp["context"], nega, num_inference_steps=3, cfg_scale=5.0,
                        torch_dtype=torch.float32)
    finally:
        P.model_fn_wan_video = orig
    with torch.no_grad():
        ref = O.denoise_loop(lambda latents, timestep, context: O.model_fn_wan_video(sd, cfg, latents, timestep, context),
                             inp["latents"], 3, 5.0, torch.float32, dict(context=inp["context"]), dict(context=nega))
    assert O.parity_metrics(got, ref)["max_abs"] <= 1e-4


def test_product_path_fails_loudly_without_cuda():
    """No CPU fallback: CPU tensors (or a missing extension) raise instead of silently computing elsewhere."""
    x = torch.randn(8, 256)
    with pytest.raises(_lib.WvdError):
        ops.ln_modulate(x, eps=1e-6)
    with pytest.raises(_lib.WvdError):
        ops.linear(x.bfloat16(), torch.randn(256, 256).bfloat16())
    with pytest.raises(_lib.WvdError):
        ops.attention(x.bfloat16(), x.bfloat16(), x.bfloat16(), 2)
    fix = dict(size="tiny", seeds=dict(dit=0, vace=3, lora=2, inputs=1), perturb=True, weight_scale=1.0,
               with_vace=False, lora=False, latent_shape=(1, 16, 3, 8, 12), timestep=700.0)
    dit, _ = build_models(fix)
    with pytest.raises(_lib.WvdError):
        run_model_fn(fix, dit, None, ops)          # default backend on CPU tensors


def test_unsupported_surface_raises():
    fix = dict(size="tiny", seeds=dict(dit=0, vace=3, lora=2, inputs=1), perturb=True, weight_scale=1.0,
               with_vace=False, lora=False, latent_shape=(1, 16, 3, 8, 12), timestep=700.0)
    dit, _ = build_models(fix)
    with pytest.raises(NotImplementedError):
        run_model_fn(fix, dit, None, cpu_backend, clip_feature=torch.zeros(1))
    with pytest.raises(NotImplementedError):
        V.WanModel(has_image_input=True, **O.DIT_CONFIGS["tiny"])


def test_lora_loader_matches_the_reference_merge():
    """video_styler_b200.GeneralLoRALoader (mirror of diffsynth/lora/__init__.py:4-45) against the oracle's merge, which
    tests/golden pins to the REAL loader (tiny_vace_lora was generated with it): same targets found through
    named_modules(), same W + alpha * B @ A bit for bit, adapter-named and 'diffusion_model.'-prefixed keys included."""
    import video_styler_b200 as V
    vcfg = O.VACE_CONFIGS["tiny"]
    vsd = O.make_state_dict(O.vace_param_shapes(vcfg), seed=3, perturb_norms=True)
    lsd = O.make_lora_state_dict(vcfg, seed=2, rank=16)
    vace = V.VaceWanModel(has_image_input=False, **vcfg)
    vace.load_state_dict({k: v.clone() for k, v in vsd.items()}, strict=True, assign=True)
    n = V.GeneralLoRALoader(device="cpu", torch_dtype=torch.float32).load(vace, lsd, alpha=0.7)
    want = {k: v.clone() for k, v in vsd.items()}
    O.lora_merge(want, lsd, alpha=0.7)
    assert n == sum(1 for k in lsd if ".lora_B." in k) > 0
    got = vace.state_dict()
    changed = 0
    for k, v in want.items():
        assert torch.equal(got[k], v), k
        changed += int(not torch.equal(v, vsd[k]))
    assert changed == n
    # key forms of get_name_dict: adapter name optional, 'diffusion_model.' prefix dropped
    nd = V.GeneralLoRALoader().get_name_dict({"diffusion_model.vace_blocks.0.ffn.0.lora_B.weight": 0, "vace_blocks.1.self_attn.q.lora_B.default.weight": 0})
    assert set(nd) == {"vace_blocks.0.ffn.0", "vace_blocks.1.self_attn.q"}


def test_block_with_reference_complex_freqs_and_padded_rope_rows(golden_dir):
    """engine.as_rope_info: the reference's complex (N, 1, 64) freqs drive the block through the per-token RoPE mode
    (same result as the table + grid mode); rows past the grid (padding of the last Ulysses shard) are tolerated."""
    fix = _load(golden_dir, "tiny_vace_lora")
    dit, _ = build_models(fix)
    cfg = O.DIT_CONFIGS["tiny"]
    inp = O.make_inputs(fix["latent_shape"], cfg["text_dim"], seed=1, with_vace=True)
    with torch.no_grad():
        ctx = dit.text_embedding(inp["context"])[0]
        x = fix["block0_in"][0].clone()
        ws = engine.workspace(x.shape[0], dit.dim, cfg["ffn_dim"], ctx.shape[0], x.dtype, x.device)
        rope = engine.as_rope_info(O.rope_freqs(128, 3, 4, 6), "cpu", cpu_backend)
        assert tuple(rope.grid) == (0, 0, 0) and rope.table.shape == (72, 64, 2)
        y = engine.dit_block_forward(dit.blocks[0], x, ctx, fix["t_mod"], rope, ws, cpu_backend)
    assert O.parity_metrics(y.unsqueeze(0), fix["block0_out"])["max_abs"] <= 5e-5
    with pytest.raises(TypeError):
        engine.as_rope_info(torch.zeros(72, 1, 64), "cpu", cpu_backend)
    # padded rows: 80 rows on a 72-token grid starting at offset 0 -> rows 72..79 take the last frame's angles
    tab = cpu_backend.make_rope_table(O.rope_tables_3d(128), "cpu")
    q = torch.randn(80, 256)
    w = torch.ones(256)
    out, _ = cpu_backend.qk_rmsnorm_rope(q.clone(), None, w, None, 1e-6, tab, (3, 4, 6), 0)
    ref, _ = cpu_backend.qk_rmsnorm_rope(q[:72].clone(), None, w, None, 1e-6, tab, (3, 4, 6), 0)
    assert torch.equal(out[:72], ref) and torch.isfinite(out).all()


def test_text_cache_is_bit_identical_and_invalidates(golden_dir):
    """engine.TextCache (text embedding + cross-attention K/V once per prompt): same bits as the uncached call, a hit on
    the second call, a miss when the context tensor is modified in place or a cross-attention weight changes."""
    fix = _load(golden_dir, "tiny_vace_lora")
    dit, vace = build_models(fix)
    base = run_model_fn(fix, dit, vace, cpu_backend)
    cfg = O.DIT_CONFIGS["tiny"]
    inp = O.make_inputs(fix["latent_shape"], cfg["text_dim"], seed=1, with_vace=True)
    cache = V.TextCache()
    ts = torch.tensor([fix["timestep"]])

    def call(ctx):
        with torch.no_grad():
            return V.model_fn_wan_video(dit=dit, vace=vace, latents=inp["latents"], timestep=ts, context=ctx,
                                        vace_context=inp["vace_context"], vace_scale=1.0, ops=cpu_backend, text_cache=cache)
    ctx = inp["context"].clone()
    a, b = call(ctx), call(ctx)
    assert torch.equal(a, base) and torch.equal(b, base) and (cache.hits, cache.misses) == (1, 1)
    n_kv = len(cache.entry(ctx)["kv"])
    assert n_kv == len(dit.blocks) + len(vace.vace_blocks)
    ctx.mul_(0.5)                                              # in-place edit: version counter moves -> recomputed
    c = call(ctx)
    assert not torch.equal(c, base) and cache.misses == 2
    with torch.no_grad():
        dit.blocks[1].cross_attn.k.weight.add_(0.01)           # e.g. a LoRA merged after the first call
        want = V.model_fn_wan_video(dit=dit, vace=vace, latents=inp["latents"], timestep=ts, context=ctx,
                                    vace_context=inp["vace_context"], vace_scale=1.0, ops=cpu_backend)
    assert torch.equal(call(ctx), want)


def test_denoise_switches_are_bit_identical():
    """denoise(cache_text / fused_step on or off) gives the same latents (CPU stand-in kernels; the fused step itself is
    checked against the eager expressions on the GPU)."""
    f = dict(size="tiny", seeds=dict(dit=0, vace=3, lora=2, inputs=1), perturb=True, weight_scale=1.0, with_vace=False, lora=False)
    dit, _ = build_models(f)
    inp = O.make_inputs((1, 16, 3, 8, 12), O.DIT_CONFIGS["tiny"]["text_dim"], seed=1)
    nega = torch.zeros_like(inp["context"])
    import functools
    import video_styler_b200.pipeline as P
    orig = P.model_fn_wan_video
    P.model_fn_wan_video = functools.partial(orig, ops=cpu_backend)
    try:
        outs = [V.denoise(dit, None, inp["latents"], inp["context"], nega, num_inference_steps=2, cfg_scale=5.0,
                          torch_dtype=torch.float32, cache_text=c) for c in (False, True)]
    finally:
        P.model_fn_wan_video = orig
    assert torch.equal(outs[0], outs[1])


def test_engine_sends_v_first_when_the_exchange_asks_for_it(golden_dir):
    """Ulysses fast path, host side: an exchange that exposes ``v_ready`` gets the v projection FIRST (its scatter then
    overlaps the q | k projection); the block output is unchanged.  (The real peer-memory exchange needs GPUs; here a
    recording stand-in checks the call order and that v is final when announced.)"""
    fix = _load(golden_dir, "tiny_t2v")
    dit, _ = build_models(fix)
    cfg = O.DIT_CONFIGS[fix["size"]]
    blk = dit.blocks[0]
    g = torch.Generator().manual_seed(0)
    n, d = 96, cfg["dim"]
    x = torch.randn(n, d, generator=g)
    ctx = torch.randn(16, d, generator=g)
    t_mod = torch.randn(1, 6, d, generator=g)
    rope = engine.RopeInfo(cpu_backend.make_rope_table(dit.freqs, "cpu"), (2, 6, 8))
    ws = engine.workspace(n, d, cfg["ffn_dim"], 16, torch.float32, "cpu")
    want = engine.dit_block_forward(blk, x.clone(), ctx, t_mod, rope, ws, cpu_backend).clone()

    class Recording(engine.SelfAttnExchange):
        world = 2                      # > 1: take the v-first branch (attention itself stays local in this stand-in)
        events = []

        def v_ready(self, ops, qkv, heads):
            self.events.append(("v_ready", qkv[:, 2 * heads * 128:].clone()))

        def norm_rope_attend(self, ops, qkv, heads, wq, wk, eps, rope, out, ws):
            self.events.append(("attend", qkv[:, 2 * heads * 128:].clone()))
            return super().norm_rope_attend(ops, qkv, heads, wq, wk, eps, rope, out, ws)
    ex = Recording()
    got = engine.dit_block_forward(blk, x.clone(), ctx, t_mod, rope, ws, cpu_backend, ex)
    assert [e[0] for e in ex.events] == ["v_ready", "attend"]
    assert torch.equal(ex.events[0][1], ex.events[1][1])          # v was final when it was announced
    assert torch.allclose(got, want, rtol=0, atol=1e-6)
