"""Developer tool: one launch each of cuBLAS (F.linear), the wvd 1-CTA and the wvd CTA-pair GEMM at two c3 shapes, for
`ncu --set full -k regex:"gemm_bf16_kernel|nvjet|cutlass|gemm"`.  Warm-up launches come first (skip them with -s)."""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_styler_b200 import _lib, ops  # noqa: E402

DEV = "cuda"
g = torch.Generator(device=DEV).manual_seed(0)
shapes = [(29640, 5120, 5120), (29640, 13824, 5120)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (m, n, k) in shapes:
    x = torch.randn(m, k, device=DEV, generator=g).bfloat16()
    w = (torch.randn(n, k, device=DEV, generator=g) / math.sqrt(k)).bfloat16()
    b = torch.randn(n, device=DEV, generator=g).bfloat16()
    out = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    for rep in range(2):
        F.linear(x, w, b)
        ops.linear(x, w, b, out=out, variant=_lib.GEMM_1CTA)
        ops.linear(x, w, b, out=out, variant=_lib.GEMM_2CTA)
    torch.cuda.synchronize()
print("ok", _lib.debug_flags())
