"""Developer tool: per-phase cycle counters of the softmax warps and the MMA issuer of one CTA pair of the cta_group::2
attention kernel.  Needs a profiling build of the library:
    WVD_NVCC_FLAGS=-DWVD_ATTN_PROF python -m video_styler_b200.build --force && cp video_styler_b200/libwvd.so video_styler_b200/libwvd_prof.so
    python tools/attn_prof_cg2.py --lib video_styler_b200/libwvd_prof.so"""
import argparse, ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_styler_b200 import _lib, ops
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default="video_styler_b200/libwvd_prof.so")
ap.add_argument("--tokens", type=int, default=29640)
ap.add_argument("--heads", type=int, default=40)
a = ap.parse_args()
_lib.LIB_PATH = os.path.abspath(a.lib)
n, h = a.tokens, a.heads
d = h * 128
q = torch.randn(n, 3 * d, device="cuda").bfloat16()
out = torch.empty(n, d, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.attention(q[:, :d], q[:, d:2 * d], q[:, 2 * d:], h, out=out, kernel=_lib.ATTN_CG2)
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
fn = _lib.load().wvd_debug_attention_profile
fn.argtypes, fn.restype = [ctypes.c_void_p], ctypes.c_int
fn(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.attention(q[:, :d], q[:, d:2 * d], q[:, 2 * d:], h, out=out, kernel=_lib.ATTN_CG2)
e1.record()
torch.cuda.synchronize()
fn(None)
b = buf.cpu().tolist()
print(f"kernel {e0.elapsed_time(e1):.3f} ms  ({n} tokens, {h} heads)")
for cta in range(2):
    for w in (0, 4):
        o = b[(cta * 8 + w) * 8:(cta * 8 + w + 1) * 8]
        it = max(o[6], 1)
        print(f"CTA {cta} softmax warp {w} (warpgroup {w // 4}): per own tile: wait_S {o[0]/it:7.1f} ld {o[1]/it:6.1f} rowmax {o[2]/it:6.1f} "
              f"m-handoff {o[3]/it:6.1f} decide {o[4]/it:6.1f} exp+handover {o[5]/it:7.1f}  total {sum(o[:6])/it:7.1f} over {o[6]} tiles")

# timeline of 32 KV steps of the leader CTA (clock ticks relative to the first event)
ev = b[256:256 + 32 * 16]
t0 = min(v for v in ev if v > 0)
names = {0: "sm.begin", 1: "sm.S_ready", 2: "sm.m_pub", 3: "sm.handover0", 4: "sm.handover1"}
print("step  " + "  ".join(f"{names[e]:>14s}" for e in sorted(names)))
for j in range(32):
    row = ev[j * 16:(j + 1) * 16]
    print(f"{40 + j:4d}  " + "  ".join(f"{((row[e] - t0) & 0xffffffff) if row[e] else -1:14d}" for e in sorted(names)))
