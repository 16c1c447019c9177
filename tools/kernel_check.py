"""Developer diagnostic: run every wvd kernel against torch references on the GPU and print error statistics and
timings (never asserts; the pytest suite under tests/ holds the pass/fail gates).

    python tools/kernel_check.py [--only gemm,attn,ew] [--big]
"""
import argparse
import math
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_styler_b200 import _lib, ops  # noqa: E402

dev = "cuda"


def stats(name, got, ref, extra=""):
    g, r = got.double().flatten(), ref.double().flatten()
    diff = (g - r).abs()
    rel = float((g - r).norm() / r.norm().clamp_min(1e-30))
    mism = float((got != ref).double().mean()) if got.dtype == ref.dtype else float("nan")
    nan = int(torch.isnan(got.float()).sum())
    print(f"  {name:46s} relL2 {rel:9.3e}  max|d| {float(diff.max()):9.3e}  mismatch {mism:8.2e}  nan {nan} {extra}", flush=True)
    return rel


def timeit(fn, iters=10, warm=3):
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


def flags(tag):
    f = _lib.debug_flags()
    if f["timeouts"]:
        print(f"  !! watchdog after {tag}: {f}", flush=True)
    return f["timeouts"]


def check_ew(big):
    print("== elementwise ==")
    for (n, d) in [(72, 256), (1280, 1536), (29640 if big else 4096, 5120)]:
        for dt in (torch.bfloat16, torch.float32):
            g = torch.Generator(device=dev).manual_seed(0)
            x = torch.randn(n, d, device=dev, generator=g).to(dt) * 2 + 0.3
            sh = (torch.randn(d, device=dev, generator=g) * 0.5).to(dt)
            sc = (torch.randn(d, device=dev, generator=g) * 0.5).to(dt)
            w = (1 + 0.1 * torch.randn(d, device=dev, generator=g)).to(dt)
            b = (0.1 * torch.randn(d, device=dev, generator=g)).to(dt)
            ref = F.layer_norm(x, (d,), None, None, 1e-6) * (1 + sc) + sh
            stats(f"ln_modulate {n}x{d} {str(dt)[6:]}", ops.ln_modulate(x, sh, sc, eps=1e-6), ref)
            ref = F.layer_norm(x, (d,), w, b, 1e-6)
            stats(f"ln_affine   {n}x{d} {str(dt)[6:]}", ops.ln_modulate(x, weight=w, bias=b, eps=1e-6), ref)
            # rmsnorm + rope
            heads = d // 128
            gf, gh, gw = 3, 4, n // 12 if n % 12 == 0 else 1
            if gf * gh * gw != n:
                gf, gh, gw = 1, 1, n
            if gw > 1024:
                gf, gh, gw = n // (30 * 52) if n % (30 * 52) == 0 else 1, 30, 52
                if gf * gh * gw != n:
                    gf, gh, gw = 4, n // 4 // 64, 64
            ok_grid = gf * gh * gw == n and gh <= 1024 and gw <= 1024 and gf <= 1024
            q = torch.randn(n, d, device=dev, generator=g).to(dt)
            k = torch.randn(n, d, device=dev, generator=g).to(dt)
            wq = (1 + 0.1 * torch.randn(d, device=dev, generator=g)).to(dt)
            wk = (1 + 0.1 * torch.randn(d, device=dev, generator=g)).to(dt)

            def rms(t, w_):
                tf = t.float()
                return (tf * torch.rsqrt(tf.pow(2).mean(-1, keepdim=True) + 1e-6)).to(t.dtype) * w_

            def rope(t, fr):
                tc = torch.view_as_complex(t.to(torch.float64).reshape(n, heads, -1, 2))
                return torch.view_as_real(tc * fr).flatten(1).to(t.dtype)
            if ok_grid:
                sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
                from oracle import wan_oracle as O
                tabs = O.rope_tables_3d(128)
                fr = O.rope_freqs(128, gf, gh, gw).to(dev)
                table = ops.make_rope_table(tabs, dev)
                qo, ko = ops.qk_rmsnorm_rope(q.clone(), k.clone(), wq, wk, 1e-6, table, (gf, gh, gw))
                stats(f"rms_rope q  {n}x{d} {str(dt)[6:]} grid {gf}x{gh}x{gw}", qo, rope(rms(q, wq), fr))
                stats(f"rms_rope k  {n}x{d} {str(dt)[6:]}", ko, rope(rms(k, wk), fr))
            qo, _ = ops.qk_rmsnorm_rope(q.clone(), None, wq, None, 1e-6)
            stats(f"rmsnorm     {n}x{d} {str(dt)[6:]}", qo, rms(q, wq))
            y = torch.randn(n, d, device=dev, generator=g).to(dt)
            stats(f"scale_add   {n}x{d} {str(dt)[6:]}", ops.scale_add(x, y, 0.75), x + y * 0.75)
            stats(f"gate_resid  {n}x{d} {str(dt)[6:]}", ops.gate_residual(x, sc, y), x + sc * y)
    n, d = (29640 if big else 4096), 5120
    x = torch.randn(n, d, device=dev).bfloat16()
    sh = torch.randn(d, device=dev).bfloat16()
    o = torch.empty_like(x)
    ms = timeit(lambda: ops.ln_modulate(x, sh, sh, eps=1e-6, out=o))
    print(f"  ln_modulate {n}x{d} bf16: {ms*1e3:.1f} us  -> {2*n*d*2/ms/1e6:.0f} GB/s")
    k = torch.randn(n, d, device=dev).bfloat16()
    tabs = None
    from oracle import wan_oracle as O
    table = ops.make_rope_table(O.rope_tables_3d(128), dev)
    grid = (n // (30 * 52), 30, 52) if n % 1560 == 0 else (4, n // 256, 64)
    ms = timeit(lambda: ops.qk_rmsnorm_rope(x, k, sh, sh, 1e-6, table, grid))
    print(f"  qk_rmsnorm_rope {n}x{d} bf16: {ms*1e3:.1f} us  -> {4*n*d*2/ms/1e6:.0f} GB/s")


def gemm_ref(x, w, b, epi, gate, res):
    y = (x.float() @ w.float().t())
    if b is not None:
        y = y + b.float()
    y = y.to(x.dtype)
    if epi == ops.EPI_BIAS_GELU:
        y = F.gelu(y, approximate="tanh")
    if epi == ops.EPI_BIAS_GATE_RES:
        y = res + gate * y
    if epi == ops.EPI_BIAS_RES:
        y = res + y
    return y


def check_gemm(big):
    print("== gemm (tcgen05) ==")
    torch.backends.cuda.matmul.allow_tf32 = False
    # exact-integer single tile: any layout/descriptor error shows up as large integer differences
    for (m, n, k) in [(128, 256, 64), (128, 256, 256), (256, 512, 128)]:
        g = torch.Generator(device=dev).manual_seed(1)
        x = torch.randint(-2, 3, (m, k), device=dev, generator=g).bfloat16()
        w = torch.randint(-2, 3, (n, k), device=dev, generator=g).bfloat16()
        out = ops.linear(x, w)
        ref = (x.float() @ w.float().t()).bfloat16()
        rel = stats(f"gemm int {m}x{n}x{k}", out, ref)
        if flags("gemm int") or rel > 1e-3:
            bad = (out != ref)
            rows = bad.any(1).nonzero().flatten()[:8].tolist()
            cols = bad.any(0).nonzero().flatten()[:8].tolist()
            print("    first bad rows", rows, "cols", cols)
            print("    got [0,:8]", out[0, :8].tolist(), "ref", ref[0, :8].tolist())
            bm = bad.view(m // 32, 32, n // 32, 32).any(3).any(1).int()
            print("    32x32 block error map:\n", bm.cpu().numpy())
    shapes = [(72, 256, 256), (200, 512, 320), (1280, 1536, 1536), (1000, 8960, 1536), (520, 64, 512)]
    if big:
        shapes += [(29640, 5120, 5120), (29640, 13824, 5120), (29640, 5120, 13824), (512, 5120, 5120)]
    else:
        shapes += [(4096, 5120, 5120)]
    for (m, n, k) in shapes:
        g = torch.Generator(device=dev).manual_seed(2)
        x = torch.randn(m, k, device=dev, generator=g).bfloat16()
        w = (torch.randn(n, k, device=dev, generator=g) / math.sqrt(k)).bfloat16()
        b = (torch.randn(n, device=dev, generator=g) * 0.1).bfloat16()
        gate = torch.randn(n, device=dev, generator=g).bfloat16()
        res = torch.randn(m, n, device=dev, generator=g).bfloat16()
        for epi, nm in [(ops.EPI_BIAS, "bias"), (ops.EPI_BIAS_GELU, "gelu"), (ops.EPI_BIAS_RES, "res"), (ops.EPI_BIAS_GATE_RES, "gate")]:
            out = ops.linear(x, w, b, epi, gate if epi == ops.EPI_BIAS_GATE_RES else None,
                             res if epi in (ops.EPI_BIAS_RES, ops.EPI_BIAS_GATE_RES) else None)
            stats(f"gemm {nm:5s} {m}x{n}x{k}", out, gemm_ref(x, w, b, epi, gate, res))
            if flags(f"gemm {nm}"):
                return
        # strided output / in-place residual
        buf = torch.zeros(m, 3 * n, device=dev, dtype=torch.bfloat16)
        ops.linear(x, w, b, out=buf[:, n:2 * n])
        stats(f"gemm strided-out {m}x{n}x{k}", buf[:, n:2 * n], gemm_ref(x, w, b, 0, None, None),
              extra=f"untouched={float(buf[:, :n].abs().max())+float(buf[:, 2*n:].abs().max())}")
        r2 = res.clone()
        ops.linear(x, w, b, ops.EPI_BIAS_GATE_RES, gate, r2, out=r2)
        stats(f"gemm in-place gate {m}x{n}x{k}", r2, gemm_ref(x, w, b, ops.EPI_BIAS_GATE_RES, gate, res))
        if m * n * k > 1e10:
            o = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
            ms = timeit(lambda: ops.linear(x, w, b, out=o), iters=5)
            ms_t = timeit(lambda: F.linear(x, w, b), iters=5)
            print(f"  time {m}x{n}x{k}: wvd {ms:.3f} ms = {2*m*n*k/ms/1e9:.0f} TFLOPS | torch/cuBLAS {ms_t:.3f} ms = {2*m*n*k/ms_t/1e9:.0f} TFLOPS", flush=True)


def check_attn(big):
    print("== attention (tcgen05) ==")
    cases = [(128, 128, 1), (256, 256, 2), (72, 72, 2), (300, 512, 2), (1280, 1280, 12), (1000, 777, 3)]
    if big:
        cases += [(29640, 29640, 40), (29640, 512, 40)]
    else:
        cases += [(4096, 4096, 8)]
    for (sq, sk, h) in cases:
        g = torch.Generator(device=dev).manual_seed(3)
        qkv = torch.randn(max(sq, sk), 3 * h * 128, device=dev, generator=g).bfloat16()
        q, k, v = qkv[:sq, :h * 128], qkv[:sk, h * 128:2 * h * 128], qkv[:sk, 2 * h * 128:]
        out = ops.attention(q, k, v, h)
        bad = flags(f"attn {sq}x{sk}x{h}")

        def ref_fn(dt):
            qh = q.to(dt).view(sq, h, 128).transpose(0, 1)[None]
            kh = k.to(dt).view(sk, h, 128).transpose(0, 1)[None]
            vh = v.to(dt).view(sk, h, 128).transpose(0, 1)[None]
            return F.scaled_dot_product_attention(qh, kh, vh)[0].transpose(0, 1).reshape(sq, h * 128)
        if sq * sk * h < 4e8:
            stats(f"attn {sq}x{sk} h{h} vs fp32 sdpa", out.float(), ref_fn(torch.float32))
        refb = ref_fn(torch.bfloat16)
        stats(f"attn {sq}x{sk} h{h} vs bf16 sdpa", out, refb)
        if bad:
            print("    out[0,:8]", out[0, :8].tolist(), "ref", refb[0, :8].tolist())
            return
        if sq * sk * h > 1e8:
            o = torch.empty(sq, h * 128, device=dev, dtype=torch.bfloat16)
            ms = timeit(lambda: ops.attention(q, k, v, h, out=o), iters=3, warm=1)
            ms_t = timeit(lambda: ref_fn(torch.bfloat16), iters=3, warm=1)
            fl = 4.0 * sq * sk * h * 128
            print(f"  time attn {sq}x{sk} h{h}: wvd {ms:.3f} ms = {fl/ms/1e9:.0f} TFLOPS | torch sdpa {ms_t:.3f} ms = {fl/ms_t/1e9:.0f} TFLOPS", flush=True)
            try:
                from flash_attn import flash_attn_func
                qf, kf, vf = (t.reshape(1, -1, h, 128).contiguous() for t in (q, k, v))
                ms_f = timeit(lambda: flash_attn_func(qf, kf, vf), iters=3, warm=1)
                print(f"       flash_attn2 {ms_f:.3f} ms = {fl/ms_f/1e9:.0f} TFLOPS", flush=True)
            except Exception as e:  # noqa
                print("       flash_attn2 unavailable:", repr(e)[:100])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="ew,gemm,attn")
    ap.add_argument("--big", action="store_true")
    a = ap.parse_args()
    print(torch.cuda.get_device_name(0), "lib version", _lib.load().wvd_version(), flush=True)
    t0 = time.time()
    for part in a.only.split(","):
        try:
            {"ew": check_ew, "gemm": check_gemm, "attn": check_attn}[part](a.big)
        except Exception as e:  # noqa
            import traceback
            traceback.print_exc()
            print(f"!! {part} raised {e!r}", flush=True)
    print(f"done in {time.time()-t0:.1f}s")
