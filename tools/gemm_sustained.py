"""Sustained A/B of the GEMM contenders under the power cap: each contender runs back to back for `--secs` seconds
(hundreds of launches, steady-state clocks), in rotating order, at the shapes of the c3 step.  The step is power-capped
(DESIGN.md section 4), so a burst measurement right after an idle gap flatters whoever runs first."""
import argparse
import math
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_styler_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--secs", type=float, default=0.6)
ap.add_argument("--rounds", type=int, default=2)
ap.add_argument("--shapes", default="29640x5120x5120,29640x13824x5120,29640x5120x13824,3705x5120x5120")
ap.add_argument("--out", default="gpurun_out/gemm_sustained.txt")
ap.add_argument("--bands", default="", help="developer: comma list of rasterisation band widths to sweep on the 2-CTA / 1-CTA variants")
a = ap.parse_args()
DEV = "cuda"
g = torch.Generator(device=DEV).manual_seed(0)
lines = ["# sustained (%.1f s per contender, rotating order, %d rounds) TFLOP/s on one B200, bias epilogue" % (a.secs, a.rounds)]
for shp in a.shapes.split(","):
    m, n, k = (int(v) for v in shp.split("x"))
    x = torch.randn(m, k, device=DEV, generator=g).bfloat16()
    w = (torch.randn(n, k, device=DEV, generator=g) / math.sqrt(k)).bfloat16()
    b = torch.randn(n, device=DEV, generator=g).bfloat16()
    out = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    cont = [("cublas", lambda: F.linear(x, w, b)),
            ("wvd-1cta", lambda: ops.linear(x, w, b, out=out, variant=_lib.GEMM_1CTA)),
            ("wvd-2cta", lambda: ops.linear(x, w, b, out=out, variant=_lib.GEMM_2CTA)),
            ("wvd-2cta-m512", lambda: ops.linear(x, w, b, out=out, variant=_lib.GEMM_2CTA_M512)),
            ("wvd-auto", lambda: ops.linear(x, w, b, out=out))]
    for bd in [int(v) for v in a.bands.split(",") if v]:
        cont.append((f"2cta-b{bd}", lambda bd=bd: ops.linear(x, w, b, out=out, variant=_lib.GEMM_2CTA | (bd << 8))))
        cont.append((f"1cta-b{bd}", lambda bd=bd: ops.linear(x, w, b, out=out, variant=_lib.GEMM_1CTA | (bd << 8))))
    fl = 2.0 * m * n * k
    res = {nm: [] for nm, _ in cont}
    for r in range(a.rounds):
        order = cont[r % len(cont):] + cont[:r % len(cont)]
        for nm, fn in order:
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            cnt = 0
            while time.perf_counter() - t0 < a.secs:
                for _ in range(20):
                    fn()
                cnt += 20
                torch.cuda.synchronize()       # keeps the launch queue short; 20 launches >> sync latency
            e1.record(); torch.cuda.synchronize()
            res[nm].append(fl * cnt / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    lines.append("  %-22s " % shp + "   ".join("%s %s" % (nm, "/".join("%.0f" % v for v in vs)) for nm, vs in res.items()))
    print(lines[-1], flush=True)
    del x, w, b, out
os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
open(a.out, "w").write("\n".join(lines) + "\n")
