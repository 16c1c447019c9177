"""Golden fixtures of the umT5 encoder and the keyframe-editor arithmetic from the REAL reference (build container
only)  --  TEST INFRASTRUCTURE.      python oracle/make_golden_aux.py      # writes tests/golden/{t5_tiny,t5_small,editor_step,vae_tiling}.pt

Loads the synthetic state dict with strict=True into the real ``WanTextEncoder`` (pins the key names), runs its forward
in fp32 on CPU; calls the real ``WanVideoEditorPipeline`` methods (compute_velocity_correction, construct_rope_ids,
prepare_coupled_noise, compute_metrics) and the real ``FlowMatchScheduler.step`` on seeded inputs.
"""
import importlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import aux_oracle as A, ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    w, _, _ = ref_shim.load()
    R = importlib.import_module("diffsynth.models.wan_video_text_encoder")
    E = importlib.import_module("diffsynth.pipelines.wan_video_editor")
    os.makedirs(OUT, exist_ok=True)
    for name, length, valid in (("tiny", 40, 23), ("small", 96, 61)):
        cfg = A.T5_CONFIGS[name]
        sd = A.make_t5_state_dict(cfg, seed=0)
        enc = R.WanTextEncoder(**cfg).eval()
        res = enc.load_state_dict(sd, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        ids, mask = A.make_t5_inputs(cfg, length, valid, seed=1)
        with torch.no_grad():
            out = enc(ids, mask)
            out_nomask = enc(ids)
        torch.save(dict(config=name, length=length, valid=valid, seeds=dict(weights=0, inputs=1), output=out.clone(),
                        output_nomask=out_nomask.clone(), torch_version=torch.__version__), os.path.join(OUT, f"t5_{name}.pt"))
        print(f"t5_{name}: out {tuple(out.shape)} |x| {float(out.abs().mean()):.4f}")

    pipe = object.__new__(E.WanVideoEditorPipeline)          # the methods under test use no instance state
    keys = (0, 3, 6)
    inp = A.make_editor_inputs(keyframes=keys, seed=5)
    t, k = inp["z_main"].shape[2], len(keys)
    cases = {}
    for tag, (alpha, beta, dt) in dict(default=(10.0, 0.0, 19.53125), beta=(4.0, 0.5, 7.25), last=(10.0, 0.0, 0.0)).items():
        v = inp["v_nega"] + 5.0 * (inp["v_posi"] - inp["v_nega"])
        v_main, v_edit = torch.split(v, [t, k], dim=2)
        vm, ve = pipe.compute_velocity_correction(inp["z_main"], inp["z_edit"], v_main, v_edit, list(keys), dt, alpha, beta)
        sch = w.FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True)
        sch.set_timesteps(50, denoising_strength=1.0, shift=5.0)
        ts = sch.timesteps[10]
        zm = sch.step(vm, ts, inp["z_main"])
        ze = sch.step(ve, ts, inp["z_edit"])
        metrics = pipe.compute_metrics(inp["z_main"], inp["z_edit"], v_main, v_edit, list(keys), dt)
        cases[tag] = dict(alpha=alpha, beta=beta, dt=dt, v_main_corrected=vm.clone(), v_edit_corrected=ve.clone(),
                          z_main_next=zm.clone(), z_edit_next=ze.clone(), metrics=metrics,
                          dsigma=float(sch.sigmas[11] - sch.sigmas[10]), step_index=10)
    noise_main, noise_edit = pipe.prepare_coupled_noise((1, 16, 7, 8, 12), list(keys), seed=7, device="cpu")
    rope_ids = pipe.construct_rope_ids(7, list(keys), device="cpu")
    torch.save(dict(keyframes=keys, shape=(1, 16, 7, 8, 12), seed=5, cfg_scale=5.0, cases=cases, noise_main=noise_main,
                    noise_edit=noise_edit, noise_seed=7, rope_ids=rope_ids, torch_version=torch.__version__),
               os.path.join(OUT, "editor_step.pt"))
    print("editor_step:", {k_: float(v_["z_main_next"].abs().mean()) for k_, v_ in cases.items()})




def vae_goldens():
    """The REAL WanVideoVAE tiling methods (tiled_decode / tiled_encode, on the CPU as the reference runs them) around the
    stand-in model, fp32 and bf16."""
    ref_shim.load()
    V = importlib.import_module("diffsynth.models.wan_video_vae")
    vae = object.__new__(V.WanVideoVAE)
    torch.nn.Module.__init__(vae)
    vae.model, vae.upsampling_factor, vae.z_dim = A.ToyVAEModel(), 8, 16
    vae.scale = [torch.zeros(16), torch.ones(16)]
    out = {}
    for name, (mode, shape, size, stride) in A.VAE_CASES.items():
        for dt in (torch.float32, torch.bfloat16):
            src = A.make_vae_source(shape, dtype=dt)
            with torch.no_grad():
                res = vae.tiled_decode(src, "cpu", size, stride) if mode == "decode" else vae.tiled_encode(src, "cpu", size, stride)
            out[(name, str(dt))] = res.clone()
        print(f"vae {name}: {tuple(res.shape)} |x| {float(res.float().abs().mean()):.4f}")
    torch.save(dict(cases=out, seed=11, torch_version=torch.__version__), os.path.join(OUT, "vae_tiling.pt"))


if __name__ == "__main__":
    main()
    vae_goldens()
