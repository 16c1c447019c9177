"""Generate golden fixtures by running the REAL reference (build container only)  --  TEST INFRASTRUCTURE.

    python oracle/make_golden.py            # writes tests/golden/*.pt

For each case it (1) draws the synthetic state dicts with ``wan_oracle.make_state_dict``, (2) loads them with
``strict=True`` into the real ``WanModel`` / ``VaceWanModel`` from /root/reference (this pins the state-dict
key names), (3) optionally merges the LoRA stand-in with the real ``GeneralLoRALoader``, (4) runs the real
``model_fn_wan_video`` on CPU/fp32 and (5) stores the seeds, inputs and outputs.  Fixtures hold only small
tensors (the weights are regenerated from the seed by the tests).
"""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim, wan_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def build_reference(dit_mod, vace_mod, cfg, vcfg, sd, vsd):
    with torch.no_grad():
        dit = dit_mod.WanModel(has_image_input=False, **cfg)
        missing = dit.load_state_dict(sd, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        vace = None
        if vcfg is not None:
            vace = vace_mod.VaceWanModel(has_image_input=False, **vcfg)
            vace.load_state_dict(vsd, strict=True)
    return dit.eval(), (vace.eval() if vace is not None else None)


def run_case(name, size, latent_shape, with_vace, lora, timestep, w, dit_mod, vace_mod, perturb=True,
             store_inter=True, weight_scale=1.0):
    cfg = O.DIT_CONFIGS[size]
    vcfg = O.VACE_CONFIGS[size] if with_vace else None
    sd = O.make_state_dict(O.dit_param_shapes(cfg), seed=0, perturb_norms=perturb, weight_scale=weight_scale)
    vsd = O.make_state_dict(O.vace_param_shapes(vcfg), seed=3, perturb_norms=perturb,
                            weight_scale=weight_scale) if with_vace else None
    dit, vace = build_reference(dit_mod, vace_mod, cfg, vcfg, sd, vsd)
    if lora:
        lsd = O.make_lora_state_dict(vcfg, seed=2, rank=16)
        from diffsynth.lora import GeneralLoRALoader
        GeneralLoRALoader(device="cpu", torch_dtype=torch.float32).load(vace, lsd, alpha=1.0)
    inp = O.make_inputs(latent_shape, cfg["text_dim"], seed=1, with_vace=with_vace)
    ts = torch.tensor([timestep], dtype=torch.float32)
    t0 = time.time()
    with torch.no_grad():
        out = w.model_fn_wan_video(dit, vace=vace, latents=inp["latents"], timestep=ts, context=inp["context"],
                                   vace_context=inp.get("vace_context"), vace_scale=1.0)
    dt = time.time() - t0
    fix = dict(name=name, size=size, latent_shape=tuple(latent_shape), with_vace=with_vace, lora=lora,
               lora_rank=16, timestep=timestep, perturb=perturb, weight_scale=weight_scale,
               seeds=dict(dit=0, vace=3, lora=2, inputs=1), output=out.clone(), ref_seconds=dt,
               torch_version=torch.__version__)
    if store_inter:
        # per-block intermediates from the real reference blocks (block-level parity)
        with torch.no_grad():
            from einops import rearrange
            t = dit.time_embedding(dit_mod.sinusoidal_embedding_1d(dit.freq_dim, ts))
            t_mod = dit.time_projection(t).unflatten(1, (6, dit.dim))
            ctx = dit.text_embedding(inp["context"])
            x = dit.patchify(inp["latents"])
            f, h, ww = x.shape[2:]
            x = rearrange(x, 'b c f h w -> b (f h w) c').contiguous()
            freqs = torch.cat([dit.freqs[0][:f].view(f, 1, 1, -1).expand(f, h, ww, -1),
                               dit.freqs[1][:h].view(1, h, 1, -1).expand(f, h, ww, -1),
                               dit.freqs[2][:ww].view(1, 1, ww, -1).expand(f, h, ww, -1)],
                              dim=-1).reshape(f * h * ww, 1, -1)
            fix["t_mod"] = t_mod.clone()
            fix["block0_in"] = x.clone()
            fix["block0_out"] = dit.blocks[0](x, ctx, t_mod, freqs).clone()
            sa = dit.blocks[0].self_attn
            q = sa.norm_q(sa.q(x))
            fix["rope_q"] = dit_mod.rope_apply(q, freqs, sa.num_heads).clone()
            if vace is not None:
                fix["hints"] = [h_.clone() for h_ in vace(x, inp["vace_context"], ctx, t_mod, freqs)]
    torch.save(fix, os.path.join(OUT, name + ".pt"))
    print(f"{name}: out {tuple(out.shape)} ref {dt:.2f}s  |out| {out.abs().mean():.4f}")


def main():
    assert ref_shim.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    w, dit_mod, vace_mod = ref_shim.load()
    run_case("tiny_t2v", "tiny", (1, 16, 3, 8, 12), False, False, 1000.0, w, dit_mod, vace_mod)
    run_case("tiny_vace_lora", "tiny", (1, 16, 3, 8, 12), True, True, 832.0, w, dit_mod, vace_mod)
    run_case("small_vace", "small", (1, 16, 5, 16, 16), True, False, 502.0, w, dit_mod, vace_mod)
    # scheduler fixture (flow_match.py) and the bf16 timestep rounding of wan_video_new.py:526
    from diffsynth.schedulers.flow_match import FlowMatchScheduler
    sch = FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True)
    sch.set_timesteps(50, shift=5.0)
    x = torch.randn(1, 16, 3, 8, 12, generator=torch.Generator().manual_seed(5))
    v = torch.randn(1, 16, 3, 8, 12, generator=torch.Generator().manual_seed(6))
    steps = {i: sch.step(v, sch.timesteps[i], x).clone() for i in (0, 25, 49)}
    torch.save(dict(sigmas=sch.sigmas.clone(), timesteps=sch.timesteps.clone(), x=x, v=v, steps=steps,
                    ts_bf16=sch.timesteps.to(torch.bfloat16).float()), os.path.join(OUT, "flow_match.pt"))
    print("flow_match: ok")
    if os.environ.get("WVD_GOLDEN_C1", "1") == "1":
        # config c1 of BASELINE.json: 1.3B, 17 frames 256x256 -> latent (1,16,5,32,32), fp32, default init
        run_case("c1_1p3B", "1.3B", (1, 16, 5, 32, 32), False, False, 1000.0, w, dit_mod, vace_mod,
                 perturb=False, store_inter=False)


if __name__ == "__main__":
    main()
