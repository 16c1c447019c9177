"""The oracle (oracle/wan_oracle.py) against golden vectors produced by the REAL reference
(oracle/make_golden.py, run in the build container).  CPU only."""
import os

import pytest
import torch

from oracle import wan_oracle as O


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)


def _run_oracle(fix, return_intermediates=False):
    cfg = O.DIT_CONFIGS[fix["size"]]
    vcfg = O.VACE_CONFIGS[fix["size"]] if fix["with_vace"] else None
    sd = O.make_state_dict(O.dit_param_shapes(cfg), seed=fix["seeds"]["dit"], perturb_norms=fix["perturb"],
                           weight_scale=fix["weight_scale"])
    vsd = None
    if vcfg is not None:
        vsd = O.make_state_dict(O.vace_param_shapes(vcfg), seed=fix["seeds"]["vace"], perturb_norms=fix["perturb"],
                                weight_scale=fix["weight_scale"])
        if fix["lora"]:
            lsd = O.make_lora_state_dict(vcfg, seed=fix["seeds"]["lora"], rank=fix["lora_rank"])
            assert O.lora_merge(vsd, lsd, alpha=1.0) == 10 * len(vcfg["vace_layers"])
    inp = O.make_inputs(fix["latent_shape"], cfg["text_dim"], seed=fix["seeds"]["inputs"], with_vace=fix["with_vace"])
    ts = torch.tensor([fix["timestep"]], dtype=torch.float32)
    with torch.no_grad():
        return O.model_fn_wan_video(sd, cfg, inp["latents"], ts, inp["context"], vsd, vcfg,
                                    inp.get("vace_context"), 1.0, return_intermediates=return_intermediates)


@pytest.mark.parametrize("name", ["tiny_t2v", "tiny_vace_lora", "small_vace"])
def test_oracle_matches_reference_golden(golden_dir, name):
    fix = _load(golden_dir, name)
    out, inter = _run_oracle(fix, return_intermediates=True)
    m = O.parity_metrics(out, fix["output"])
    assert m["max_abs"] <= 2e-5 and m["rel_l2"] <= 1e-5, m
    if fix["with_vace"]:
        for a, b in zip(inter["hints"], fix["hints"]):
            assert O.parity_metrics(a, b)["max_abs"] <= 2e-5


def test_oracle_block_and_rope_intermediates(golden_dir):
    fix = _load(golden_dir, "tiny_t2v")
    cfg = O.DIT_CONFIGS["tiny"]
    sd = O.make_state_dict(O.dit_param_shapes(cfg), seed=0, perturb_norms=True)
    inp = O.make_inputs(fix["latent_shape"], cfg["text_dim"], seed=1)
    ctx = O._lin(sd, "text_embedding.2", torch.nn.functional.gelu(O._lin(sd, "text_embedding.0", inp["context"]),
                                                                  approximate="tanh"))
    f, h, w = 3, 4, 6
    freqs = O.rope_freqs(128, f, h, w)
    x = fix["block0_in"]
    y = O.dit_block(sd, "blocks.0.", x, ctx, fix["t_mod"], freqs, cfg["num_heads"], cfg["eps"])
    assert O.parity_metrics(y, fix["block0_out"])["max_abs"] <= 1e-5
    q = O.rms_norm(O._lin(sd, "blocks.0.self_attn.q", x), sd["blocks.0.self_attn.norm_q.weight"], cfg["eps"])
    assert O.parity_metrics(O.rope_apply(q, freqs, 2), fix["rope_q"])["max_abs"] <= 1e-6


def test_oracle_rope_closed_form():
    """pair j<22 -> frame, 22<=j<43 -> row, 43<=j<64 -> column; theta = pos*10000^(-2j'/dim_axis)."""
    f, h, w = 3, 4, 5
    fr = O.rope_freqs(128, f, h, w)[:, 0]             # (N, 64) complex128
    n = torch.arange(f * h * w)
    pos = [n // (h * w), (n // w) % h, n % w]
    dims, offs = [44, 42, 42], [0, 22, 43]
    for ax in range(3):
        half = dims[ax] // 2
        j = torch.arange(half, dtype=torch.float64)
        ang = pos[ax].double()[:, None] * (10000.0 ** (-2 * j / dims[ax]))[None]
        got = fr[:, offs[ax]:offs[ax] + half]
        assert (got.real - torch.cos(ang)).abs().max() < 1e-12 and (got.imag - torch.sin(ang)).abs().max() < 1e-12


def test_oracle_scheduler(golden_dir):
    fix = _load(golden_dir, "flow_match")
    sigmas, ts = O.flow_match_schedule(50, 5.0)
    assert torch.equal(sigmas, fix["sigmas"]) and torch.equal(ts, fix["timesteps"])
    for i, ref in fix["steps"].items():
        got = O.flow_match_step(sigmas, ts, fix["v"], ts[i], fix["x"])
        assert torch.equal(got, ref)
    # wan_video_new.py:526 rounds the timestep to the model dtype: 937.5->936 etc.
    assert torch.equal(ts.to(torch.bfloat16).float(), fix["ts_bf16"])
    assert float(ts.to(torch.bfloat16)[0]) == 1000.0


@pytest.mark.slow
def test_oracle_c1_matches_reference(golden_dir):
    """BASELINE config c1: Wan2.1-T2V-1.3B random-init, latent (1,16,5,32,32), fp32, one call."""
    fix = _load(golden_dir, "c1_1p3B")
    out = _run_oracle(fix)
    m = O.parity_metrics(out, fix["output"])
    assert m["max_abs"] <= 5e-5 and m["rel_l2"] <= 2e-5, m
