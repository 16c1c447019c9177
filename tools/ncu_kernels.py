"""Developer tool: one launch sequence of every hot kernel at the c3 shapes (for `ncu --set full` captures).

    ncu --set full --clock-control none -k regex:'gemm_bf16|ln_modulate|qk_rmsnorm' -s 8 -c 8 -o out python tools/ncu_kernels.py
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_styler_b200 import ops
from oracle import wan_oracle as O

n, d, f = 29640, 5120, 13824
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(n, d, device="cuda", generator=g).bfloat16()
wq = (torch.randn(d, d, device="cuda", generator=g) / 72).bfloat16()
w1 = (torch.randn(f, d, device="cuda", generator=g) / 72).bfloat16()
w2 = (torch.randn(d, f, device="cuda", generator=g) / 118).bfloat16()
b_d = torch.zeros(d, device="cuda", dtype=torch.bfloat16)
b_f = torch.zeros(f, device="cuda", dtype=torch.bfloat16)
gate = torch.ones(d, device="cuda", dtype=torch.bfloat16)
sh = torch.zeros(d, device="cuda", dtype=torch.bfloat16)
qkv = torch.randn(n, 3 * d, device="cuda", generator=g).bfloat16()
mid = torch.empty(n, f, device="cuda", dtype=torch.bfloat16)
h = torch.empty(n, d, device="cuda", dtype=torch.bfloat16)
table = ops.make_rope_table(O.rope_tables_3d(128), "cuda")
for rep in range(2):      # first pass = warm-up (skipped with ncu -s)
    ops.ln_modulate(x, sh, gate, eps=1e-6, out=h)                                   # LayerNorm + modulate
    ops.linear(h, wq, b_d, out=qkv[:, :d])                                          # q projection (EPI_BIAS)
    ops.qk_rmsnorm_rope(qkv[:, :d], qkv[:, d:2 * d], gate, gate, 1e-6, table, (19, 30, 52), 0)
    ops.linear(h, w1, b_f, ops.EPI_BIAS_GELU, out=mid)                              # FFN up + GELU-tanh
    ops.linear(mid, w2, b_d, ops.EPI_BIAS_GATE_RES, gate=gate, residual=x, out=x)   # FFN down + gate + residual
torch.cuda.synchronize()
print("ok")
