"""SURVEY section 8(f)3-4 on the B200, through the C ABI: the umT5 encoder (attention-with-bias kernel, the
elementwise-product GEMM epilogue, the whole encoder) and the keyframe-editor step kernel / loop, against the aux oracle
and the golden vectors of the REAL reference.  fp32 mode <= 1e-4, bf16 cos >= 0.999 and relL2 <= 1e-2 (BASELINE.json);
the editor kernel keeps the reference's rounding points and is held to bit-identity with the same torch ops on the GPU."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import aux_oracle as A, wan_oracle as O  # noqa: E402
from video_styler_b200 import _lib, ops, wan_video_editor as E, wan_video_text_encoder as T  # noqa: E402

DEV = "cuda"


def assert_bf16_parity(ours, ref_bf16, ref_fp32, what):
    """BASELINE.json's bf16 bar -- cos >= 0.999 and relL2 <= 1e-2 against the reference's bf16 path -- with the noise-floor
    protocol of SURVEY.md section 7 where the reference's OWN bf16 rounding noise exceeds that budget (24 umT5 layers, a
    CFG-amplified multi-step loop): two bf16 evaluations that differ only in accumulation order are then ~sqrt(2) noise
    floors apart, so the bar becomes "no further from the fp32 truth than the reference's bf16 path is" (+10 %)."""
    m, floor, ours32 = O.parity_metrics(ours, ref_bf16), O.parity_metrics(ref_bf16, ref_fp32), O.parity_metrics(ours, ref_fp32)
    print(f"{what}: wvd vs oracle-bf16 {m}; oracle-bf16 vs oracle-fp32 (noise floor) {floor}; wvd vs oracle-fp32 {ours32}")
    assert m["cos"] >= 0.999, m
    assert m["rel_l2"] <= 1e-2 or ours32["rel_l2"] <= 1.1 * floor["rel_l2"] + 1e-3, (m, floor, ours32)
    return m, floor, ours32


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)


def _ref_attention_bias(q, k, v, h, bias, mask, dtype):
    """The reference's T5Attention core (wan_video_text_encoder.py:68-84) on (L, H*64) inputs in `dtype`."""
    lq, lk = q.shape[0], k.shape[0]
    qh, kh, vh = (t.to(dtype).view(1, t.shape[0], h, 64) for t in (q, k, v))
    attn_bias = qh.new_zeros(1, h, lq, lk)
    if bias is not None:
        idx = (torch.arange(lk, device=q.device)[None, :] - torch.arange(lq, device=q.device)[:, None]) + (lq - 1)
        attn_bias += bias.to(dtype)[:, idx].unsqueeze(0)
    if mask is not None:
        attn_bias.masked_fill_(mask.view(1, 1, 1, -1) == 0, torch.finfo(dtype).min)
    attn = torch.einsum("binc,bjnc->bnij", qh, kh) + attn_bias
    attn = F.softmax(attn.float(), dim=-1).type_as(attn)
    return torch.einsum("bnij,bjnc->binc", attn, vh).reshape(lq, h * 64)


@pytest.mark.parametrize("lq,lk,h,valid", [(1, 1, 1, 1), (40, 40, 4, 23), (33, 70, 2, 70), (512, 512, 8, 77), (100, 130, 3, 0)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 4e-3)])
def test_attention_bias_kernel_matches_the_reference_expression(lq, lk, h, valid, dtype, tol):
    g = torch.Generator(device=DEV).manual_seed(lq * 7 + lk)
    q, k, v = (torch.randn(n, h * 64, device=DEV, generator=g).to(dtype) for n in (lq, lk, lk))
    q = q * 0.35                                            # T5 does not scale: keep the logits in a softmax-relevant range
    bias = (torch.randn(h, lq + lk - 1, device=DEV, generator=g) * 0.7).to(dtype)
    mask = None
    if valid:
        mask = torch.zeros(lk, dtype=torch.int32, device=DEV)
        mask[:valid] = 1
    out = ops.attention_bias(q, k, v, h, bias, mask, 1.0)
    ref = _ref_attention_bias(q, k, v, h, bias, mask, dtype)
    m = O.parity_metrics(out, ref)
    assert m["rel_l2"] <= tol, m
    # strided views (the encoder's fused q|k|v buffer) and no bias / no mask
    if lq == lk:
        buf = torch.cat([q, k, v], dim=1).contiguous()
        d = h * 64
        out2 = ops.attention_bias(buf[:, :d], buf[:, d:2 * d], buf[:, 2 * d:], h, None, None, 1.0)
        m2 = O.parity_metrics(out2, _ref_attention_bias(q, k, v, h, None, None, dtype))
        assert m2["rel_l2"] <= tol, m2


def test_attention_bias_argument_errors():
    q = torch.zeros(8, 128, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(_lib.WvdError):
        ops.attention_bias(q, q, q, 1)                                            # head_dim 128
    with pytest.raises(_lib.WvdError):
        ops.attention_bias(q, q, q, 2, bias=torch.zeros(2, 10, device=DEV, dtype=torch.bfloat16))   # table too short
    with pytest.raises(_lib.WvdError):
        ops.attention_bias(q.cpu(), q.cpu(), q.cpu(), 2)


@pytest.mark.parametrize("m,n,k", [(40, 640, 256), (512, 10240, 4096), (200, 1280, 512)])
def test_gemm_elementwise_product_epilogue(m, n, k):
    g = torch.Generator(device=DEV).manual_seed(m + n)
    x = torch.randn(m, k, device=DEV, generator=g).bfloat16()
    w = (torch.randn(n, k, device=DEV, generator=g) / math.sqrt(k)).bfloat16()
    other = torch.randn(m, n, device=DEV, generator=g).bfloat16()
    out = ops.linear(x, w, None, ops.EPI_BIAS_MUL, residual=other)
    ref = F.linear(x, w) * other                                                  # bf16 GEMM output rounded, then the bf16 product
    m_ = O.parity_metrics(out, ref)
    assert m_["rel_l2"] <= 2e-3, m_
    xf, wf, of = x.float(), w.float(), other.float()
    outf = ops.linear(xf, wf, None, ops.EPI_BIAS_MUL, residual=of)
    assert O.parity_metrics(outf, F.linear(xf, wf) * of)["rel_l2"] <= 1e-5


def _encoder(cfg, sd, dtype):
    enc = T.WanTextEncoder(**cfg).eval().requires_grad_(False)
    res = enc.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return enc.to(device=DEV, dtype=dtype)


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_text_encoder_fp32_matches_the_real_reference(golden_dir, name):
    fix = _load(golden_dir, f"t5_{name}")
    cfg = A.T5_CONFIGS[name]
    enc = _encoder(cfg, A.make_t5_state_dict(cfg, seed=0), torch.float32)
    ids, mask = A.make_t5_inputs(cfg, fix["length"], fix["valid"], seed=1)
    with torch.no_grad():
        out = enc(ids.to(DEV), mask.to(DEV))
        out2 = enc(ids.to(DEV), None)
    m, m2 = O.parity_metrics(out, fix["output"]), O.parity_metrics(out2, fix["output_nomask"])
    print(f"umT5 {name} fp32 vs the real reference: masked {m}, unmasked {m2}")
    assert m["rel_l2"] <= 1e-4 and m2["rel_l2"] <= 1e-4, (m, m2)
    assert _lib.debug_flags()["timeouts"] == 0


@pytest.mark.parametrize("name,length,valid,layers", [("small", 96, 61, None), ("umt5-xxl", 512, 77, 2), ("umt5-xxl", 512, 300, 24)])
def test_text_encoder_bf16_matches_the_oracle_in_bf16(name, length, valid, layers):
    """bf16 on the GPU against the oracle restatement in bf16 on the same device with the same bf16 weights -- at the
    real umT5-XXL width (dim 4096, 64 heads, ffn 10240, 512 tokens), 2 layers and the full 24."""
    cfg = dict(A.T5_CONFIGS[name])
    if layers is not None:
        cfg["num_layers"] = layers
    if name == "umt5-xxl":
        cfg["vocab"] = 4096                                  # the embedding table is a gather, not arithmetic: keep it small
    sd = A.make_t5_state_dict(cfg, seed=0, dtype=torch.bfloat16, device=DEV)
    enc = _encoder(cfg, sd, torch.bfloat16)
    ids, mask = A.make_t5_inputs(cfg, length, valid, seed=1)
    ids, mask = ids.to(DEV), mask.to(DEV)
    with torch.no_grad():
        out = T.encode_prompt(enc, ids, mask)
        ref = A.encode_prompt(sd, cfg, ids, mask)
        ref32 = A.encode_prompt({k: v.float() for k, v in sd.items()}, cfg, ids, mask)
    assert_bf16_parity(out, ref, ref32, f"umT5 {name} x{cfg['num_layers']} bf16")
    assert float(out[:, valid:].abs().max()) == 0.0
    assert _lib.debug_flags()["timeouts"] == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_editor_step_kernel(golden_dir, dtype):
    """fp32 against the REAL reference's outputs; bf16 bit-identical to the reference's op sequence run with torch on the
    same GPU (same rounding points); joint and split velocity layouts; corrected-velocity mode; no-CFG mode."""
    fix = _load(golden_dir, "editor_step")
    keys = list(fix["keyframes"])
    inp = {k: v.to(DEV).to(dtype) for k, v in A.make_editor_inputs(fix["shape"], keys, seed=fix["seed"]).items()}
    t = fix["shape"][2]
    km = E.KeyframeMap(t, keys, DEV)
    for tag, c in fix["cases"].items():
        zm, ze = ops.editor_step(inp["z_main"], inp["z_edit"], inp["v_posi"], inp["v_nega"], km.frame_to_key, km.key_idx,
                                 fix["cfg_scale"], c["dt"], c["alpha"], c["beta"], c["dsigma"], euler=True)
        rm, re_ = A.editor_step(inp["z_main"], inp["z_edit"], inp["v_posi"], inp["v_nega"], keys, fix["cfg_scale"], c["dt"],
                                c["alpha"], c["beta"], c["dsigma"])
        if dtype == torch.float32:
            assert O.parity_metrics(zm, c["z_main_next"])["rel_l2"] <= 1e-6 and O.parity_metrics(ze, c["z_edit_next"])["rel_l2"] <= 1e-6, tag
        else:
            assert torch.equal(zm, rm) and torch.equal(ze, re_), tag
        v = inp["v_nega"] + fix["cfg_scale"] * (inp["v_posi"] - inp["v_nega"])
        vm, ve = E.compute_velocity_correction(inp["z_main"], inp["z_edit"], v[:, :, :t], v[:, :, t:], keys, c["dt"], c["alpha"], c["beta"])
        rvm, rve = A.velocity_correction(inp["z_main"], inp["z_edit"], v[:, :, :t], v[:, :, t:], keys, c["dt"], c["alpha"], c["beta"])
        if dtype == torch.float32:
            assert O.parity_metrics(vm, c["v_main_corrected"])["rel_l2"] <= 1e-6 and O.parity_metrics(ve, c["v_edit_corrected"])["rel_l2"] <= 1e-6
        else:
            assert torch.equal(vm, rvm) and torch.equal(ve, rve), tag
    zm, ze = ops.editor_step(inp["z_main"], inp["z_edit"], inp["v_posi"], None, km.frame_to_key, km.key_idx, 1.0, 3.0, 2.0, 0.0, -0.02)
    rm, re_ = A.editor_step(inp["z_main"], inp["z_edit"], inp["v_posi"], None, keys, 1.0, 3.0, 2.0, 0.0, -0.02)
    assert O.parity_metrics(zm, rm)["rel_l2"] <= 1e-6 and O.parity_metrics(ze, re_)["rel_l2"] <= 1e-6


def test_editor_step_at_the_c3_latent_size():
    """(1, 16, 19, 60, 104) main latents + 5 keyframes, bf16: bit-identical to the torch op sequence."""
    g = torch.Generator(device=DEV).manual_seed(9)
    keys = [0, 4, 9, 14, 18]
    r = lambda *s: torch.randn(*s, device=DEV, generator=g).bfloat16()      # noqa: E731
    zm, ze, vp, vn = r(1, 16, 19, 60, 104), r(1, 16, 5, 60, 104), r(1, 16, 24, 60, 104), r(1, 16, 24, 60, 104)
    km = E.KeyframeMap(19, keys, DEV)
    a, b = ops.editor_step(zm, ze, vp, vn, km.frame_to_key, km.key_idx, 5.0, 19.5, 10.0, 0.3, -0.0123)
    ra, rb = A.editor_step(zm, ze, vp, vn, keys, 5.0, 19.5, 10.0, 0.3, -0.0123)
    assert torch.equal(a, ra) and torch.equal(b, rb)


def test_edit_denoise_on_the_gpu_matches_the_reference_loop():
    """3 editing steps of the tiny DiT in bf16: edit_denoise (joint DiT call with rope ids [0..T-1 | keyframes] on the wvd
    kernels + the fused step kernel) against the reference loop restated with the oracle on the same GPU."""
    import video_styler_b200 as V
    cfg = O.DIT_CONFIGS["tiny"]
    sd = {k: v.to(DEV).bfloat16() for k, v in O.make_state_dict(O.dit_param_shapes(cfg), seed=0, perturb_norms=True).items()}
    dit = V.WanModel(has_image_input=False, **cfg).eval().requires_grad_(False)
    dit.load_state_dict(sd, strict=True, assign=True)
    keys = [0, 2]
    g = torch.Generator(device=DEV).manual_seed(3)
    z_main = torch.randn(1, 16, 3, 8, 12, device=DEV, generator=g).bfloat16()
    z_edit = z_main[:, :, keys].clone()                                     # coupled noise
    ctx_p = torch.randn(1, 12, cfg["text_dim"], device=DEV, generator=g).bfloat16()
    ctx_n = torch.zeros_like(ctx_p)
    steps = 3
    with torch.no_grad():
        zm, ze = E.edit_denoise(dit, z_main, z_edit, ctx_p, ctx_n, keys, num_inference_steps=steps, cfg_scale=5.0, alpha=2.0, beta=0.0)
        sch = V.FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True)
        sch.set_timesteps(steps, shift=5.0)
        rope = torch.tensor([0, 1, 2] + keys)                               # the oracle gathers from its CPU tables

        def reference_loop(dtype):
            w = {k: v.to(dtype) for k, v in sd.items()}
            rm, re_ = z_main.to(dtype), z_edit.to(dtype)
            for i, ts in enumerate(sch.timesteps):
                tsb = ts.unsqueeze(0).to(dtype=torch.bfloat16, device=DEV).to(dtype)      # the pipeline's bf16-rounded timestep
                zc = torch.cat([rm, re_], dim=2)
                vp = O.model_fn_wan_video(w, cfg, zc, tsb, ctx_p.to(dtype), rope_indices=rope)
                vn = O.model_fn_wan_video(w, cfg, zc, tsb, ctx_n.to(dtype), rope_indices=rope)
                dt = float(sch.timesteps[i] - sch.timesteps[i + 1]) if i < steps - 1 else 0.0
                rm, re_ = A.editor_step(rm, re_, vp, vn, keys, 5.0, dt, 2.0, 0.0, sch.dsigma(ts))
            return rm, re_
        rm, re_ = reference_loop(torch.bfloat16)
        fm, fe = reference_loop(torch.float32)
    assert_bf16_parity(zm, rm, fm, "edit_denoise z_main, 3 steps, bf16")
    assert_bf16_parity(ze, re_, fe, "edit_denoise z_edit, 3 steps, bf16")
    assert _lib.debug_flags()["timeouts"] == 0


# ---- VAE tiling layer (the tiling part): blend + finalize kernels ----
@pytest.mark.parametrize("name", ["decode_small", "decode_ragged", "decode_one_tile", "encode_small", "encode_odd"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_vae_tiling_kernels_match_the_real_reference(golden_dir, name, dtype):
    """TiledVAE on the GPU (wvd_tile_blend / wvd_tile_finalize; vector and scalar paths) against the outputs of the REAL
    WanVideoVAE.tiled_decode / tiled_encode: bit-identical in bf16 AND fp32 (every operation is a single rounding)."""
    from video_styler_b200 import wan_video_vae as VA
    fix = _load(golden_dir, "vae_tiling")
    mode, shape, size, stride = A.VAE_CASES[name]
    src = A.make_vae_source(shape, seed=fix["seed"], dtype=dtype).to(DEV)
    gold = fix["cases"][(name, str(dtype))]
    t = VA.TiledVAE(A.ToyVAEModel())
    with torch.no_grad():
        out = t.tiled_decode(src, DEV, size, stride) if mode == "decode" else t.tiled_encode(src, DEV, size, stride)
        ref = A.vae_tiled(A.ToyVAEModel(), src, size, stride, mode)            # the same op sequence with torch on the GPU
    assert torch.equal(out, ref)
    m = O.parity_metrics(out, gold)            # the toy model's own arithmetic may differ CPU vs GPU by an ulp (avg_pool / linspace)
    assert m["rel_l2"] <= (1e-6 if dtype == torch.float32 else 4e-3), m


def test_vae_tiling_at_the_c3_video_size():
    """Decode blending at the real size: (1, 16, 19, 60, 104) latents -> (1, 3, 73, 480, 832), tiles (30, 52) / (15, 26) as
    infer_ditto.py passes them; bf16, bit-identical to the reference's op sequence on the GPU; encode likewise."""
    from video_styler_b200 import wan_video_vae as VA
    t = VA.TiledVAE(A.ToyVAEModel())
    z = A.make_vae_source((1, 16, 19, 60, 104), seed=3, dtype=torch.bfloat16).to(DEV)
    with torch.no_grad():
        out = t.decode([z[0]], DEV, tiled=True, tile_size=(30, 52), tile_stride=(15, 26))
        ref = A.vae_tiled(A.ToyVAEModel(), z, (30, 52), (15, 26), "decode")
        assert out.shape == (1, 3, 73, 480, 832) and torch.equal(out[0], ref[0])
        v = A.make_vae_source((1, 3, 17, 480, 832), seed=4, dtype=torch.bfloat16).to(DEV)
        lat = t.encode([v[0]], DEV, tiled=True, tile_size=(30, 52), tile_stride=(15, 26))
        refl = A.vae_tiled(A.ToyVAEModel(), v, (240, 416), (120, 208), "encode")
        assert lat.shape == (1, 16, 5, 60, 104) and torch.equal(lat[0], refl[0])
