"""Measurement of the callers either side of the DiT on one B200 -- the `aux` leg of bench.py on
its own: one umT5-XXL prompt (this path vs the oracle's restatement with torch eager on the same GPU), one keyframe-editor
step at the c3 latent size (fused kernel vs the reference's torch op sequence) and the VAE tile blending at the c3 video size.

    python tools/aux_bench.py [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--json", default=None)
a = ap.parse_args()
res = bench.aux_leg(torch.device("cuda", 0), bench.peaks()["hbm"])
for k, v in res.items():
    print(k, json.dumps(v))
if a.json:
    json.dump(res, open(a.json, "w"), indent=1)
