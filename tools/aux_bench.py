"""Measurement of the two "next" rows either side of the DiT on one B200:

  umT5-XXL text encoder   24 layers, dim 4096, 64 heads, ffn 10240, one 512-token prompt, bf16: this path vs the oracle's
                          restatement run with torch eager on the same GPU (cuBLAS + eager softmax / norms: the reference's
                          kernel sequence); 4.83 TFLOP of projections + 0.10 TFLOP of attention per prompt
  keyframe editor step    the per-step arithmetic after the DiT calls at the c3 latent size (19 + 5 frames of 16 x 60 x 104):
                          the fused kernel vs the reference's torch op sequence; HBM-bound, algorithmic bytes =
                          read z (T+K) + v_posi, v_nega (T+K each) + write z (T+K) frames

    python tools/aux_bench.py [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import aux_oracle as A  # noqa: E402  (baseline leg only)
from video_styler_b200 import ops, wan_video_editor as E, wan_video_text_encoder as T  # noqa: E402

DEV = "cuda"


def timed(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    res = {}
    hbm = 6454.3
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        hbm = json.load(open(p))["hbm_gbs"]
    with torch.no_grad():
        cfg = dict(A.T5_CONFIGS["umt5-xxl"])
        cfg["vocab"] = 4096
        sd = A.make_t5_state_dict(cfg, seed=0, dtype=torch.bfloat16, device=DEV)
        enc = T.WanTextEncoder(**cfg).eval().requires_grad_(False)
        enc.load_state_dict(sd, strict=True)
        enc = enc.to(device=DEV, dtype=torch.bfloat16)
        ids, mask = A.make_t5_inputs(cfg, 512, 77, seed=1)
        ids, mask = ids.to(DEV), mask.to(DEV)
        ms = timed(lambda: T.encode_prompt(enc, ids, mask), 10)
        ms_ref = timed(lambda: A.encode_prompt(sd, cfg, ids, mask), 5)
        l, d, da, f, nl = 512, cfg["dim"], cfg["dim_attn"], cfg["dim_ffn"], cfg["num_layers"]
        flops = nl * (2 * l * (4 * d * da + 3 * d * f) + 4 * l * l * da)
        res["umt5_xxl_prompt"] = dict(wvd_ms=ms, torch_eager_ms=ms_ref, speedup=ms_ref / ms, tflop=flops / 1e12,
                                      wvd_tflops_per_s=flops / ms / 1e9,
                                      what="24 layers, dim 4096, 64 heads x 64, ffn 10240, 512 tokens (77 valid), bf16, batch 1")
        print(f"umT5-XXL prompt: wvd {ms:.2f} ms ({flops / ms / 1e9:.0f} TFLOP/s)  torch eager (oracle on the GPU) {ms_ref:.2f} ms  -> {ms_ref / ms:.2f}x")
        del enc, sd
        torch.cuda.empty_cache()

        g = torch.Generator(device=DEV).manual_seed(9)
        keys = [0, 4, 9, 14, 18]
        r = lambda *s: torch.randn(*s, device=DEV, generator=g).bfloat16()      # noqa: E731
        zm, ze, vp, vn = r(1, 16, 19, 60, 104), r(1, 16, 5, 60, 104), r(1, 16, 24, 60, 104), r(1, 16, 24, 60, 104)
        km = E.KeyframeMap(19, keys, DEV)
        us = 1e3 * timed(lambda: ops.editor_step(zm, ze, vp, vn, km.frame_to_key, km.key_idx, 5.0, 19.5, 10.0, 0.0, -0.0123), 200, 10)
        us_ref = 1e3 * timed(lambda: A.editor_step(zm, ze, vp, vn, keys, 5.0, 19.5, 10.0, 0.0, -0.0123), 50, 5)
        nbytes = 4 * vp.numel() * 2
        res["editor_step_c3"] = dict(wvd_us=us, torch_ops_us=us_ref, speedup=us_ref / us, algorithmic_bytes=nbytes,
                                     wvd_gb_per_s=nbytes / us / 1e3, hbm_peak_gb_per_s=hbm, frac_of_hbm_peak=nbytes / us / 1e3 / hbm,
                                     what="(1,16,19,60,104) main + 5 keyframes, CFG + velocity correction + Euler of both sets, bf16; "
                                          "9.6 MB of traffic: L2-resident, launch-latency bound")
        print(f"editor step (c3 latents): wvd {us:.1f} us ({nbytes / us / 1e3:.0f} GB/s)  torch op sequence {us_ref:.1f} us  -> {us_ref / us:.1f}x")
    if a.json:
        json.dump(res, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
