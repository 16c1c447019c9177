"""Developer tool: many launches of the c3 self-attention, each timed and compared with the first result (races show up
as slow launches -- the in-kernel watchdog -- or as results that differ between launches)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_styler_b200 import _lib, ops
n, h, reps = int(os.environ.get("N", 29640)), int(os.environ.get("H", 40)), int(os.environ.get("REPS", 60))
d = h * 128
qkv = torch.randn(n, 3 * d, device="cuda").bfloat16()
junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = torch.empty(n, d, device="cuda", dtype=torch.bfloat16)
ref = None
slow, bad = 0, 0
for r in range(reps):
    junk.add_(1)                       # some other kernel in between, flushes L2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], h, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if ref is None:
        ref = out.clone()
    same = bool(torch.equal(ref, out))
    if ms > 30 or not same:
        print(f"launch {r}: {ms:.1f} ms identical_to_first {same} flags {_lib.debug_flags()}", flush=True)
        slow += ms > 30
        bad += not same
print(f"{reps} launches: {slow} slow, {bad} differing; flags {_lib.debug_flags()}")
