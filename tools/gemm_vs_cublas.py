"""wvd GEMM (both variants + the grouped q|k|v launch) against cuBLAS (torch F.linear) on the same B200, at the shapes
the c3 / c4 / c5 steps actually run.  Writes the table the verdict asks for (profiles/r2_gemm_vs_cublas.txt).

    python tools/gemm_vs_cublas.py [--out gpurun_out/gemm_vs_cublas.txt] [--reps 20]

Timing: CUDA events over `reps` back-to-back launches after 3 warm-ups, alternating the contenders (A, B, A, B ...) so
that clocks / power state are shared; a 160 MB buffer is touched between groups to flush L2.  The bias epilogue is
included on both sides (F.linear with bias)."""
import argparse
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_styler_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/gemm_vs_cublas.txt")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--rounds", type=int, default=3)
a = ap.parse_args()
DEV = "cuda"
SHAPES = [(m, n, k) for m in (29640, 3705) for (n, k) in ((5120, 5120), (13824, 5120), (5120, 13824))]
SHAPES += [(512, 5120, 5120), (75600, 5120, 5120), (14820, 5120, 5120), (7410, 5120, 5120), (32760, 1536, 1536), (32760, 8960, 1536)]
junk = torch.empty(160 << 20, dtype=torch.uint8, device=DEV)


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


lines = ["# wvd GEMM vs cuBLAS (torch F.linear, bf16, bias included) on one B200; TFLOP/s, best of %d rounds of %d launches" % (a.rounds, a.reps),
         "# %-22s %10s %10s %10s %10s %10s   %s" % ("M x N x K", "cuBLAS", "wvd 1-CTA", "wvd 2-CTA", "wvd m512", "wvd auto", "auto / cuBLAS")]
g = torch.Generator(device=DEV).manual_seed(0)
for (m, n, k) in SHAPES:
    x = torch.randn(m, k, device=DEV, generator=g).bfloat16()
    w = (torch.randn(n, k, device=DEV, generator=g) / math.sqrt(k)).bfloat16()
    b = torch.randn(n, device=DEV, generator=g).bfloat16()
    out = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    fl = 2.0 * m * n * k
    cont = {"cublas": lambda: F.linear(x, w, b),
            "1cta": lambda: ops.linear(x, w, b, out=out, variant=_lib.GEMM_1CTA),
            "2cta": lambda: ops.linear(x, w, b, out=out, variant=_lib.GEMM_2CTA),
            "m512": lambda: ops.linear(x, w, b, out=out, variant=_lib.GEMM_2CTA_M512),
            "auto": lambda: ops.linear(x, w, b, out=out)}
    best = {kk: 1e9 for kk in cont}
    for _ in range(a.rounds):
        for kk, fn in cont.items():
            junk.add_(1)
            best[kk] = min(best[kk], timed(fn, a.reps))
    ref = F.linear(x, w, b)
    ok = all(float((ops.linear(x, w, b, variant=v).float() - ref.float()).norm() / ref.float().norm()) < 2e-3 for v in (1, 2, 3))
    tf = {kk: fl / (v * 1e-3) / 1e12 for kk, v in best.items()}
    lines.append("  %-22s %10.0f %10.0f %10.0f %10.0f %10.0f   %.3f %s" % (f"{m} x {n} x {k}", tf["cublas"], tf["1cta"], tf["2cta"], tf["m512"], tf["auto"],
                                                                     tf["auto"] / tf["cublas"], "" if ok else "MISMATCH"))
    print(lines[-1], flush=True)
    del x, w, b, out, ref

# the grouped q|k|v launch against three cuBLAS calls / one cuBLAS call on a concatenated (3N, K) weight
lines.append("# grouped q|k|v projections (N = 3 x 5120): ms per group of three")
for m in (29640, 3705):
    n, k = 5120, 5120
    x = torch.randn(m, k, device=DEV, generator=g).bfloat16()
    ws = [(torch.randn(n, k, device=DEV, generator=g) / math.sqrt(k)).bfloat16() for _ in range(3)]
    bs = [torch.randn(n, device=DEV, generator=g).bfloat16() for _ in range(3)]
    wcat, bcat = torch.cat(ws), torch.cat(bs)
    out = torch.empty(m, 3 * n, device=DEV, dtype=torch.bfloat16)
    cont = {"cublas 3 calls": lambda: [F.linear(x, ws[i], bs[i]) for i in range(3)],
            "cublas 1 call on cat(W)": lambda: F.linear(x, wcat, bcat),
            "wvd 3 launches": lambda: [ops.linear(x, ws[i], bs[i], out=out[:, i * n:(i + 1) * n]) for i in range(3)],
            "wvd grouped 1-CTA": lambda: ops.linear_grouped(x, ws, bs, out=out, variant=_lib.GEMM_1CTA),
            "wvd grouped 2-CTA": lambda: ops.linear_grouped(x, ws, bs, out=out, variant=_lib.GEMM_2CTA),
            "wvd grouped auto": lambda: ops.linear_grouped(x, ws, bs, out=out)}
    best = {kk: 1e9 for kk in cont}
    for _ in range(a.rounds):
        for kk, fn in cont.items():
            junk.add_(1)
            best[kk] = min(best[kk], timed(fn, max(4, a.reps // 2)))
    fl = 3 * 2.0 * m * n * k
    lines.append("  M = %-6d " % m + "  ".join("%s %.3f ms (%.0f TF/s)" % (kk, v, fl / (v * 1e-3) / 1e12) for kk, v in best.items()))
    print(lines[-1], flush=True)
    del x, ws, bs, wcat, bcat, out
print("flags", _lib.debug_flags())
os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
open(a.out, "w").write("\n".join(lines) + "\n")
