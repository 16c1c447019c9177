"""SURVEY section 8(f)3-4 without a GPU: the aux oracle (oracle/aux_oracle.py) against golden vectors from the REAL
reference (oracle/make_golden_aux.py), and the host logic of the umT5 encoder / keyframe editor modules driven through the
CPU stand-in for the kernel wrappers (state-dict key parity with the reference, block orchestration, loop, errors)."""
import os

import pytest
import torch

from oracle import aux_oracle as A, wan_oracle as O
from tests import cpu_backend
from video_styler_b200 import WvdError, wan_video_editor as E, wan_video_text_encoder as T


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_t5_oracle_matches_reference_golden(golden_dir, name):
    fix = _load(golden_dir, f"t5_{name}")
    cfg = A.T5_CONFIGS[name]
    sd = A.make_t5_state_dict(cfg, seed=fix["seeds"]["weights"])
    ids, mask = A.make_t5_inputs(cfg, fix["length"], fix["valid"], seed=fix["seeds"]["inputs"])
    with torch.no_grad():
        m = O.parity_metrics(A.t5_encoder(sd, cfg, ids, mask), fix["output"])
        m2 = O.parity_metrics(A.t5_encoder(sd, cfg, ids, None), fix["output_nomask"])
    assert m["max_abs"] <= 2e-5 and m["rel_l2"] <= 1e-5, m
    assert m2["max_abs"] <= 2e-5 and m2["rel_l2"] <= 1e-5, m2


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_text_encoder_module_tree_and_host_logic(golden_dir, name):
    """The reference's state dict loads with strict=True (same key names), and forward() through the CPU stand-in of the
    kernels reproduces the real reference's output (fp32; the stand-in evaluates GELU-tanh once, not op by op: identical
    in fp32 up to rounding)."""
    fix = _load(golden_dir, f"t5_{name}")
    cfg = A.T5_CONFIGS[name]
    enc = T.WanTextEncoder(**cfg).eval().requires_grad_(False)
    res = enc.load_state_dict(A.make_t5_state_dict(cfg, seed=0), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    ids, mask = A.make_t5_inputs(cfg, fix["length"], fix["valid"], seed=1)
    with torch.no_grad():
        out = enc(ids, mask, ops=cpu_backend)
        out2 = enc(ids, None, ops=cpu_backend)
    m, m2 = O.parity_metrics(out, fix["output"]), O.parity_metrics(out2, fix["output_nomask"])
    assert m["rel_l2"] <= 1e-5 and m2["rel_l2"] <= 1e-5, (m, m2)
    # the dense bias tensor of the reference module and the kernel's table agree
    pe = enc.blocks[0].pos_embedding
    dense = pe(7, 9)
    assert dense.shape == (1, cfg["num_heads"], 7, 9)
    assert torch.equal(dense, A.t5_pos_bias(pe.embedding.weight, 7, 9))


def test_text_encoder_fails_loudly_on_cpu_and_bad_config():
    cfg = A.T5_CONFIGS["tiny"]
    enc = T.WanTextEncoder(**cfg).eval().requires_grad_(False)
    ids, mask = A.make_t5_inputs(cfg, 16, 9)
    with torch.no_grad(), pytest.raises(WvdError):
        enc(ids, mask)                               # CPU tensors on the product path: no fallback
    with pytest.raises(ValueError):
        T.WanTextEncoder(vocab=10, dim=256, dim_attn=256, dim_ffn=512, num_heads=2, num_layers=1)     # head_dim 128


def test_editor_oracle_matches_reference_golden(golden_dir):
    fix = _load(golden_dir, "editor_step")
    inp = A.make_editor_inputs(fix["shape"], fix["keyframes"], seed=fix["seed"])
    t, k = fix["shape"][2], len(fix["keyframes"])
    for tag, c in fix["cases"].items():
        v = inp["v_nega"] + fix["cfg_scale"] * (inp["v_posi"] - inp["v_nega"])
        vm, ve = A.velocity_correction(inp["z_main"], inp["z_edit"], v[:, :, :t], v[:, :, t:], fix["keyframes"], c["dt"],
                                       c["alpha"], c["beta"])
        assert torch.equal(vm, c["v_main_corrected"]) and torch.equal(ve, c["v_edit_corrected"]), tag
        zm, ze = A.editor_step(inp["z_main"], inp["z_edit"], inp["v_posi"], inp["v_nega"], fix["keyframes"], fix["cfg_scale"],
                               c["dt"], c["alpha"], c["beta"], c["dsigma"])
        assert torch.allclose(zm, c["z_main_next"], rtol=0, atol=1e-6) and torch.allclose(ze, c["z_edit_next"], rtol=0, atol=1e-6), tag


def test_editor_helpers_match_the_reference(golden_dir):
    fix = _load(golden_dir, "editor_step")
    nm, ne = E.prepare_coupled_noise(fix["shape"], fix["keyframes"], seed=fix["noise_seed"], device="cpu")
    assert torch.equal(nm, fix["noise_main"]) and torch.equal(ne, fix["noise_edit"])
    assert torch.equal(E.construct_rope_ids(fix["shape"][2], fix["keyframes"], device="cpu"), fix["rope_ids"])
    inp = A.make_editor_inputs(fix["shape"], fix["keyframes"], seed=fix["seed"])
    t = fix["shape"][2]
    v = inp["v_nega"] + fix["cfg_scale"] * (inp["v_posi"] - inp["v_nega"])
    c = fix["cases"]["default"]
    m = E.compute_metrics(inp["z_main"], inp["z_edit"], v[:, :, :t], v[:, :, t:], fix["keyframes"], c["dt"])
    assert m == pytest.approx(c["metrics"], rel=1e-6)
    vm, ve = E.compute_velocity_correction(inp["z_main"], inp["z_edit"], v[:, :, :t], v[:, :, t:], fix["keyframes"], c["dt"],
                                           c["alpha"], c["beta"], ops=cpu_backend)
    assert torch.equal(vm, c["v_main_corrected"]) and torch.equal(ve, c["v_edit_corrected"])
    with pytest.raises(ValueError):
        E.KeyframeMap(7, [1, 1], "cpu")              # duplicates: the reference's indexed += is ill-defined
    with pytest.raises(IndexError):
        E.KeyframeMap(7, [0, 7], "cpu")
    with pytest.raises(WvdError):                    # CPU tensors on the product path: no fallback
        E.compute_velocity_correction(inp["z_main"], inp["z_edit"], v[:, :, :t], v[:, :, t:], fix["keyframes"], 1.0)


def test_edit_denoise_loop_host_logic():
    """edit_denoise through the CPU stand-in == the reference loop restated with the oracle's model_fn (joint DiT call
    with rope ids [0..T-1 | keyframes], CFG, split, correction, Euler), 3 steps, fp32."""
    import video_styler_b200 as V
    cfg = O.DIT_CONFIGS["tiny"]
    sd = O.make_state_dict(O.dit_param_shapes(cfg), seed=0, perturb_norms=True)
    dit = V.WanModel(has_image_input=False, **cfg).eval().requires_grad_(False)
    dit.load_state_dict(sd, strict=True)
    keys = [0, 2]
    g = torch.Generator().manual_seed(3)
    z_main, z_edit = torch.randn(1, 16, 3, 8, 12, generator=g), torch.randn(1, 16, 2, 8, 12, generator=g)
    ctx_p = torch.randn(1, 12, cfg["text_dim"], generator=g)
    ctx_n = torch.zeros_like(ctx_p)
    steps, cfg_scale, alpha, beta = 3, 5.0, 10.0, 0.25
    with torch.no_grad():
        zm, ze = E.edit_denoise(dit, z_main, z_edit, ctx_p, ctx_n, keys, num_inference_steps=steps, cfg_scale=cfg_scale,
                                alpha=alpha, beta=beta, torch_dtype=torch.float32, ops=cpu_backend)
        sch = V.FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True)
        sch.set_timesteps(steps, shift=5.0)
        rm, re_ = z_main, z_edit
        rope = torch.tensor([0, 1, 2] + keys)
        for i, ts in enumerate(sch.timesteps):
            zc = torch.cat([rm, re_], dim=2)
            vp = O.model_fn_wan_video(sd, cfg, zc, ts.unsqueeze(0), ctx_p, rope_indices=rope)
            vn = O.model_fn_wan_video(sd, cfg, zc, ts.unsqueeze(0), ctx_n, rope_indices=rope)
            dt = float(sch.timesteps[i] - sch.timesteps[i + 1]) if i < steps - 1 else 0.0
            rm, re_ = A.editor_step(rm, re_, vp, vn, keys, cfg_scale, dt, alpha, beta, sch.dsigma(ts))
    assert O.parity_metrics(zm, rm)["rel_l2"] <= 1e-5 and O.parity_metrics(ze, re_)["rel_l2"] <= 1e-5


# ---- VAE tiling layer (the tiling part) ----
@pytest.mark.parametrize("name", ["decode_small", "decode_ragged", "decode_one_tile", "encode_small", "encode_odd"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_vae_tiling_oracle_and_host_logic_match_the_real_reference(golden_dir, name, dtype):
    """The oracle's restatement AND the package's TiledVAE (through the CPU stand-in of the two kernels) reproduce the REAL
    WanVideoVAE.tiled_decode / tiled_encode bit for bit, fp32 and bf16 (same tile enumeration, masks, rounding)."""
    from video_styler_b200 import wan_video_vae as VA
    fix = _load(golden_dir, "vae_tiling")
    mode, shape, size, stride = A.VAE_CASES[name]
    src = A.make_vae_source(shape, seed=fix["seed"], dtype=dtype)
    gold = fix["cases"][(name, str(dtype))]
    with torch.no_grad():
        assert torch.equal(A.vae_tiled(A.ToyVAEModel(), src, size, stride, mode), gold)
        t = VA.TiledVAE(A.ToyVAEModel(), ops=cpu_backend)
        out = t.tiled_decode(src, "cpu", size, stride) if mode == "decode" else t.tiled_encode(src, "cpu", size, stride)
    assert torch.equal(out, gold)


def test_vae_masks_and_errors():
    from video_styler_b200 import wan_video_vae as VA
    t = VA.TiledVAE(A.ToyVAEModel(), ops=cpu_backend)
    data = torch.zeros(1, 3, 2, 12, 20)
    for bound in [(True, True, True, True), (False, True, True, False), (False, False, False, False)]:
        assert torch.equal(t.build_mask(data, bound, (5, 8)), A.vae_build_mask(data, bound, (5, 8)))
    assert VA.tile_tasks(9, 13, (4, 6), (2, 3)) == [(h, h + 4, w, w + 6) for h in (0, 2, 4, 6) for w in (0, 3, 6, 9)]
    with pytest.raises(WvdError):                    # CPU tensors on the product path: no fallback
        VA.TiledVAE(A.ToyVAEModel()).tiled_decode(torch.zeros(1, 16, 2, 8, 8), "cpu", (4, 4), (2, 2))
