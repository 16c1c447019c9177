"""Developer tool: every kernel once at small shapes, watchdog flags printed at the end (a quick does-everything-launch check; compute-sanitizer is closed on the GPU pool, so bounds are checked by the parity tests instead)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_styler_b200 import _lib, ops
from oracle import wan_oracle as O
g = torch.Generator(device="cuda").manual_seed(0)
n, d, h = 300, 512, 4
x = torch.randn(n, d, device="cuda", generator=g).bfloat16()
w = (torch.randn(776, d, device="cuda", generator=g) / 22).bfloat16()
b = torch.zeros(776, device="cuda", dtype=torch.bfloat16)
sh = torch.zeros(d, device="cuda", dtype=torch.bfloat16); gate = torch.ones(d, device="cuda", dtype=torch.bfloat16)
ops.ln_modulate(x, sh, gate, eps=1e-6)
ops.ln_modulate(x, weight=gate, bias=sh, eps=1e-6)
y = ops.linear(x, w, b, ops.EPI_BIAS_GELU)
w2 = (torch.randn(d, 776, device="cuda", generator=g) / 28).bfloat16()
ops.linear(y, w2, sh, ops.EPI_BIAS_GATE_RES, gate=gate, residual=x, out=x.clone())
qkv = torch.randn(n, 3 * d, device="cuda", generator=g).bfloat16()
table = ops.make_rope_table(O.rope_tables_3d(128), "cuda")
ops.qk_rmsnorm_rope(qkv[:, :d], qkv[:, d:2 * d], gate, gate, 1e-6, table, (3, 10, 10), 0)
for kern in (1, 2):
    o = ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], h, kernel=kern)
outs = [torch.zeros(150, d, dtype=torch.bfloat16, device="cuda") for _ in range(2)]
recv = [torch.zeros(300, 3 * d // 2, dtype=torch.bfloat16, device="cuda") for _ in range(2)]
ops.ulysses_scatter_qkv(qkv[:150].contiguous(), h, [t.data_ptr() for t in recv], 0)
ops.attention_scatter(recv[0][:, :256], recv[0][:, 256:512], recv[0][:, 512:], 2, [t.data_ptr() for t in outs], d, 150, 0)
torch.cuda.synchronize()
print("ok", _lib.debug_flags())
