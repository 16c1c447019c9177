#!/usr/bin/env python
"""bench.py -- seconds per denoising step of the Wan2.1-VACE-14B DiT (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c5|c1] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...   (N > 1: Ulysses)

A "step" is ONE model_fn_wan_video call (one velocity prediction, batch 1) -- the default
CFG denoising step is two of them.  Synthetic latents/text/VACE context of the named shapes, random-init weights
of the named architecture with the rank-128 Ditto-LoRA stand-in merged at load.

  value      s/step with inputs resident in HBM, CUDA events, barrier + synchronize on both sides, max over ranks
  e2e        the same through the public API with HOST (pinned) buffers: H2D of latents/context/vace_context/timestep
             and D2H of the velocity inside the timed region
  roofline   dominant kernel = self-attention (49.5 % of the FLOPs at c3): algorithmic FLOPs per launch / average
             launch duration from CUDA events recorded on the launching stream inside the timed region
  cpu_baseline / --impl reference
             a MEASURED whole call of the oracle's model_fn_wan_video at BASELINE config c1 (1.3B, 1,280 tokens, fp32) on
             all host cores; gpu_c1 = this path at the same configuration; the c3 figure on the CPU is an extrapolation
             and is labelled as such (c3_extrapolated_not_measured)
  gpu_reference
             the reference's kernel sequence on the same B200 in the same run (oracle restatement on the GPU, bf16, same
             weights: cuBLAS F.linear + eager norms + torch SDPA, and FlashAttention-2 when importable)
  parity     N > 1: the N-rank output vs the single-GPU output (untimed, full depth, worst rank)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "wan_vace_14b_dit_s_per_denoise_step_832x480x73"
NCU_ATTENTION_SUMMARY = "r2_attention_cg2p_c3.txt"       # ncu --set full of the dominant kernel (tools/ncu_summary.py)


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE self-attention launch at c3, from the committed
    `ncu --set full` capture summary (profiles/, written by tools/ncu_summary.py).  None if the summary is absent."""
    p = os.path.join(ROOT, "profiles", NCU_ATTENTION_SUMMARY)
    if not os.path.exists(p):
        return None
    tot, mult = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for ln in open(p):
        f = ln.split()
        if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(f[2].replace(",", "")) * mult.get(f[1], 1.0)
    return tot or None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], hbm=d["hbm_gbs"], src="measured")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------------------------------------------
# CPU leg: the oracle (port of the reference path, pinned to the real reference's golden vectors) on the host cores.
# MEASURED whole call: BASELINE config c1 (Wan2.1-T2V-1.3B random-init, 17 frames 256x256 = 1,280 tokens, fp32, one
# model_fn_wan_video call) -- the reference's own CPU-runnable configuration.  The c3 workload
# cannot run on the CPU in bounded time (70 GB of fp32 weights, 1.75 PFLOP): its figure is an EXTRAPOLATION from the
# measured c1 throughput and is reported only under a key that says so.
# ----------------------------------------------------------------------------------------------------------------
def cpu_model_name():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_c1_call(steps=3, warmup=1):
    import torch
    from oracle import wan_oracle as O
    from video_styler_b200 import synthetic as S
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.DIT_CONFIGS["1.3B"]
    sd = O.make_state_dict(O.dit_param_shapes(cfg), seed=0)
    inp = O.make_inputs(S.WORKLOADS["c1"]["latent"], cfg["text_dim"], seed=1)
    ts = torch.tensor([1000.0])
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.model_fn_wan_video(sd, cfg, inp["latents"], ts, inp["context"])
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return statistics.median(times), times, torch.get_num_threads()


def cpu_baseline_entry(c3_flops, steps=3, warmup=1):
    from video_styler_b200 import synthetic as S
    t, times, cores = cpu_c1_call(steps, warmup)
    c1_flops = S.model_flops("1.3B", 1280, False)
    return dict(value=t, unit="s", cores=cores, kind="port", cpu=cpu_model_name(),
                sample=f"WHOLE call, measured: oracle model_fn_wan_video at BASELINE config c1 (Wan2.1-T2V-1.3B random-init, "
                       f"1,280 tokens, 30 layers, fp32, {cores} threads), median of {len(times)} after {warmup} warm-up",
                times_s=times, tflops_per_s=c1_flops / t / 1e12,
                c3_extrapolated_not_measured=dict(
                    value=t * c3_flops / c1_flops, unit="s",
                    how=f"c1 seconds x (c3 FLOPs {c3_flops/1e12:.1f} T / c1 FLOPs {c1_flops/1e12:.2f} T); the c3 model in fp32 "
                        f"needs 70 GB and ~1 h on these cores, so it is NOT run"))


def run_reference(args, rank):
    """--impl reference: the reference path's CPU implementation (the oracle port; the reference itself cannot be
    installed on the GPU box) on the host cores.  value = a MEASURED whole c1 call; config names c1, not the wvd arm's
    c3 (which cannot run on a CPU in bounded time); the c3 extrapolation is a labelled extra."""
    if rank != 0:
        return
    from video_styler_b200 import synthetic as S
    t0 = time.perf_counter()
    entry = cpu_baseline_entry(S.model_flops("14B", 29640, True), steps=max(1, min(args.steps, 5)), warmup=max(1, min(args.warmup, 2)))
    line = dict(metric=METRIC, value=entry["value"], unit="s", n_gpus=0, steps=args.steps, warmup=args.warmup,
                ms_per_step=entry["value"] * 1e3, higher_is_better=False, scaling="strong", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=workload_name("c1", 1280), tokens=1280,
                            note="the wvd arm's default workload is c3; c3 cannot run on CPU in bounded time -- see "
                                 "cpu_baseline.c3_extrapolated_not_measured; the wvd arm reports its own c1 time as gpu_c1"),
                cpu_baseline=entry,
                e2e=dict(value=entry["value"], unit="s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                wall_s=time.perf_counter() - t0)
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# GPU reference leg: the reference's kernel sequence on the SAME B200 in the same run -- the oracle's restatement of
# model_fn_wan_video executed on the GPU in bf16 with the wvd arm's own weight tensors: cuBLAS F.linear, eager
# LayerNorm / RMSNorm / fp64 RoPE / gate kernels, and attention through (i) torch SDPA (wan_video_dit.py:55-60) and
# (ii) FlashAttention-2 (wan_video_dit.py:43-48 -- what the reference dispatches to in this image), if importable.
# A baseline leg: the oracle is never the thing measured as the product.
# ----------------------------------------------------------------------------------------------------------------
def gpu_reference_leg(dit, vace, devin, t_dev, ours_out, ours_ms, size, reps=3):
    import torch
    from oracle import wan_oracle as O
    from video_styler_b200 import synthetic as S
    cfg = dict(O.DIT_CONFIGS[size]); cfg["num_layers"] = len(dit.blocks)
    sd = dict(dit.state_dict())
    vsd = vcfg = None
    if vace is not None:
        vcfg = dict(O.VACE_CONFIGS[size]); vcfg["vace_layers"] = tuple(vace.vace_layers)
        vsd = dict(vace.state_dict())

    def call():
        return O.model_fn_wan_video(sd, cfg, devin["latents"], t_dev, devin["context"], vsd, vcfg, devin.get("vace_context"), 1.0)

    def timed(fn):
        out = fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, out

    res = {}
    with torch.no_grad():
        try:
            ms, out = timed(call)
            res["sdpa"] = dict(s_per_step=ms / 1e3, attention="torch.nn.functional.scaled_dot_product_attention (cuDNN / flash backend)",
                               parity_wvd_vs_this=O.parity_metrics(ours_out, out), wvd_speedup=ms / ours_ms)
            del out
        except Exception as e:          # noqa: BLE001 -- a baseline leg must not take the bench down
            res["sdpa"] = dict(unavailable=repr(e)[:200])
        torch.cuda.empty_cache()
        try:
            from flash_attn import flash_attn_func
            orig = O.attention

            def fa2(q, k, v, num_heads):
                b, sq, _ = q.shape
                o = flash_attn_func(q.view(b, sq, num_heads, -1), k.view(b, k.shape[1], num_heads, -1),
                                    v.view(b, v.shape[1], num_heads, -1))
                return o.reshape(b, sq, -1)
            O.attention = fa2
            try:
                ms, out = timed(call)
            finally:
                O.attention = orig
            res["fa2"] = dict(s_per_step=ms / 1e3, attention="flash_attn 2.x flash_attn_func (the reference's dispatch in this image)",
                              parity_wvd_vs_this=O.parity_metrics(ours_out, out), wvd_speedup=ms / ours_ms)
            del out
        except Exception as e:          # noqa: BLE001
            res["fa2"] = dict(unavailable=repr(e)[:200])
        torch.cuda.empty_cache()
    best = min((v["s_per_step"] for v in res.values() if "s_per_step" in v), default=None)
    res["best_s_per_step"] = best
    res["wvd_speedup_vs_best"] = (best * 1e3 / ours_ms) if best else None
    res["what"] = ("oracle restatement of model_fn_wan_video on the same GPU, bf16, the wvd arm's own weight tensors: "
                   "cuBLAS F.linear + eager norm/RoPE/gate kernels + the named attention library")
    return res


# ----------------------------------------------------------------------------------------------------------------
# The callers either side of the DiT, timed in the same run: one umT5-XXL prompt (24 layers,
# 512 tokens, bf16) on this path and as the oracle's restatement with torch eager on the same GPU (cuBLAS + eager
# softmax / norms: the reference's kernel sequence), and the keyframe editor's per-step arithmetic at the c3 latent
# size (fused kernel vs the reference's torch op sequence; HBM-bound: bytes = 4 x (T + K) frames).
# ----------------------------------------------------------------------------------------------------------------
def aux_leg(dev, hbm_gbs):
    import torch
    from oracle import aux_oracle as A
    from video_styler_b200 import ops, wan_video_editor as E, wan_video_text_encoder as T

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

    res = {}
    try:
        with torch.no_grad():
            cfg = dict(A.T5_CONFIGS["umt5-xxl"]); cfg["vocab"] = 4096          # the embedding table is a gather, kept small
            sd = A.make_t5_state_dict(cfg, seed=0, dtype=torch.bfloat16, device=dev)
            enc = T.WanTextEncoder(**cfg).eval().requires_grad_(False)
            enc.load_state_dict(sd, strict=True)
            enc = enc.to(device=dev, dtype=torch.bfloat16)
            ids, mask = (t.to(dev) for t in A.make_t5_inputs(cfg, 512, 77, seed=1))
            ms = timed(lambda: T.encode_prompt(enc, ids, mask), 10)
            ms_ref = timed(lambda: A.encode_prompt(sd, cfg, ids, mask), 5)
            l, d, da, f, nl = 512, cfg["dim"], cfg["dim_attn"], cfg["dim_ffn"], cfg["num_layers"]
            flops = nl * (2 * l * (4 * d * da + 3 * d * f) + 4 * l * l * da)
            res["umt5_xxl_prompt"] = dict(wvd_ms=ms, torch_eager_ms=ms_ref, speedup=ms_ref / ms, tflop=flops / 1e12,
                                          wvd_tflops_per_s=flops / ms / 1e9,
                                          what="24 layers, dim 4096, 64 heads x 64, ffn 10240, 512 tokens (77 valid), bf16, batch 1")
            del enc, sd
            torch.cuda.empty_cache()
            g = torch.Generator(device=dev).manual_seed(9)
            keys = [0, 4, 9, 14, 18]
            r = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()      # noqa: E731
            zm, ze, vp, vn = r(1, 16, 19, 60, 104), r(1, 16, 5, 60, 104), r(1, 16, 24, 60, 104), r(1, 16, 24, 60, 104)
            km = E.KeyframeMap(19, keys, dev)
            us = 1e3 * timed(lambda: ops.editor_step(zm, ze, vp, vn, km.frame_to_key, km.key_idx, 5.0, 19.5, 10.0, 0.0, -0.0123), 200, 10)
            us_ref = 1e3 * timed(lambda: A.editor_step(zm, ze, vp, vn, keys, 5.0, 19.5, 10.0, 0.0, -0.0123), 50, 5)
            nbytes = 4 * vp.numel() * 2
            res["editor_step_c3"] = dict(wvd_us=us, torch_ops_us=us_ref, speedup=us_ref / us, algorithmic_bytes=nbytes,
                                         wvd_gb_per_s=nbytes / us / 1e3, hbm_peak_gb_per_s=hbm_gbs,
                                         what="(1,16,19,60,104) main + 5 keyframe latents: CFG + velocity correction + Euler of both sets "
                                              "in one kernel, bf16; 9.6 MB of traffic = L2-resident and launch-latency bound")
            # VAE tile blending at the c3 video size: 9 decoded tiles of (1, 3, 73, 240, 416) into (1, 3, 73, 480, 832)
            from video_styler_b200 import wan_video_vae as VA
            tasks = VA.tile_tasks(60, 104, (30, 52), (15, 26))
            tiles = [torch.randn(1, 3, 73, 240, 416, device=dev, generator=g).bfloat16() for _ in range(3)]
            values = torch.zeros(1, 3, 73, 480, 832, device=dev, dtype=torch.bfloat16)
            weight = torch.zeros(480, 832, device=dev, dtype=torch.bfloat16)
            bounds = [(h == 0, h_ >= 60, w == 0, w_ >= 104) for h, h_, w, w_ in tasks]

            def blend_wvd():
                values.zero_(); weight.zero_()
                for i, (h, _, w, _) in enumerate(tasks):
                    ops.tile_blend(values, weight, tiles[i % 3], h * 8, w * 8, bounds[i], (120, 208))
                ops.tile_finalize(values, weight, (-1.0, 1.0))

            def blend_torch(vals, wgt, tl, dv):
                vals.zero_(); wgt.zero_()
                for i, (h, _, w, _) in enumerate(tasks):
                    t_ = tl[i % 3]
                    mask = A.vae_build_mask(t_, bounds[i], (120, 208)).to(dtype=vals.dtype, device=dv)
                    vals[:, :, :, h * 8:h * 8 + 240, w * 8:w * 8 + 416] += t_ * mask
                    wgt[:, :, :, h * 8:h * 8 + 240, w * 8:w * 8 + 416] += mask
                return (vals / wgt).clamp_(-1, 1)
            w5 = torch.zeros(1, 1, 73, 480, 832, device=dev, dtype=torch.bfloat16)
            ms_b = timed(blend_wvd, 10)
            ms_bt = timed(lambda: blend_torch(values, w5, tiles, dev), 5)
            nbytes = len(tasks) * 3 * tiles[0].numel() * 2 + 2 * values.numel() * 2
            res["vae_tile_blend_c3"] = dict(wvd_ms=ms_b, torch_ops_on_gpu_ms=ms_bt, speedup=ms_bt / ms_b, algorithmic_bytes=nbytes,
                                            wvd_gb_per_s=nbytes / ms_b / 1e6, hbm_peak_gb_per_s=hbm_gbs,
                                            what="9 tiles (1,3,73,240,416) blended into (1,3,73,480,832) + normalise + clamp, bf16; bytes = "
                                                 "per tile (read tile + read / write the window) + the final pass; the reference does "
                                                 "this on the CPU with a PCIe copy per tile (wan_video_vae.py:1118)")
    except Exception as e:          # noqa: BLE001 -- an extra leg must not take the bench down
        res["unavailable"] = repr(e)[:200]
    return res


def exchange_kind():
    from video_styler_b200 import ulysses
    kinds = {("UlyssesExchange" if getattr(e, "_nccl_only", False) else type(e).__name__) for e in ulysses._EXCHANGES.values()}
    return {"P2PUlyssesExchange": "all-to-alls fused into the pack / attention kernels over NVLink peer memory",
            "UlyssesExchange": "NCCL all_to_all_single"}.get(next(iter(kinds), ""), "none")


def workload_name(key, tokens):
    return {"c3": f"c3: Wan2.1-VACE-14B DiT + merged rank-128 Ditto-LoRA stand-in, VACE context, 73 frames 832x480 ({tokens} tokens), one model_fn_wan_video call",
            "c2": f"c2: Wan2.1-T2V-1.3B DiT bf16, 81 frames 832x480 ({tokens} tokens), one model_fn_wan_video call",
            "c5": f"c5: Wan2.1-VACE-14B DiT, 81 frames 1280x720 ({tokens} tokens), one model_fn_wan_video call",
            "c1": f"c1: Wan2.1-T2V-1.3B DiT, 17 frames 256x256 ({tokens} tokens), one model_fn_wan_video call"}[key]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="wvd", choices=["wvd", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c2", "c3", "c5"])
    ap.add_argument("--layers", type=int, default=None, help="debug: fewer layers (the JSON line is then marked invalid)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the torch-eager (cuBLAS + SDPA / FA2) leg")
    ap.add_argument("--no-loop", action="store_true", help="skip the CFG-step (denoise loop) timing, plain vs fused")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the untimed sharded-vs-unsharded check")
    ap.add_argument("--no-aux", action="store_true", help="skip the umT5 prompt / keyframe-editor step timings (SURVEY 8(f)3-4)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import video_styler_b200 as V
    from video_styler_b200 import ops, synthetic as S
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    wl = S.WORKLOADS[args.workload]
    b, c, f, h, w = wl["latent"]
    tokens = f * (h // 2) * (w // 2)
    heads = S.DIT_CONFIGS[wl["size"]]["num_heads"]
    dit, vace = S.build_models(wl["size"], wl["vace"], dev, torch.bfloat16, seed=0, lora_rank=128, num_layers=args.layers)
    host = S.make_inputs(wl["latent"], with_vace=wl["vace"], seed=1, dtype=torch.bfloat16, pin=True)
    t_host = torch.tensor([832.0], dtype=torch.bfloat16).pin_memory()       # bf16(1000 * sigma_25) (wan_video_new.py:526)
    devin = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    t_dev = t_host.to(dev)
    usp = world > 1
    full_flops = S.model_flops(wl["size"], tokens, wl["vace"], num_layers=args.layers)

    def step_resident():
        return V.model_fn_wan_video(dit=dit, vace=vace, timestep=t_dev, vace_scale=1.0,
                                    use_unified_sequence_parallel=usp, **devin)

    out_host = torch.empty((b, c, f, h, w), dtype=torch.bfloat16).pin_memory()

    def step_e2e():
        d_in = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        out = V.model_fn_wan_video(dit=dit, vace=vace, timestep=t_host.to(dev, non_blocking=True), vace_scale=1.0,
                                   use_unified_sequence_parallel=usp, **d_in)
        out_host.copy_(out, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / k

    parity = None
    with torch.no_grad():
        if usp and not args.no_parity:
            # untimed: the P-rank output against the single-GPU output of the SAME model on the SAME inputs (every rank
            # computes the unsharded call itself); Ulysses is exact, so bit-identity is expected and reported
            single = V.model_fn_wan_video(dit=dit, vace=vace, timestep=t_dev, vace_scale=1.0, **devin)
            sharded = step_resident()
            torch.cuda.synchronize()
            a, b_ = sharded.double().flatten(), single.double().flatten()
            loc = torch.tensor([float((a - b_).norm() / b_.norm()),
                                1.0 - float(torch.nn.functional.cosine_similarity(a, b_, dim=0)),
                                0.0 if torch.equal(sharded, single) else 1.0], device=dev, dtype=torch.float64)
            dist.all_reduce(loc, op=dist.ReduceOp.MAX)                 # worst rank
            parity = dict(rel_l2=float(loc[0]), cos=1.0 - float(loc[1]), bit_identical=bool(float(loc[2]) == 0.0),
                          what=f"{world}-rank Ulysses output vs the single-GPU output, full depth, worst over ranks")
            del single, sharded, a, b_
        for _ in range(args.warmup):
            step_resident()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        ops.PROFILE = {}
        launches0 = ops.LAUNCHES
        torch.cuda.nvtx.range_push("wvd_timed")       # ncu --nvtx --nvtx-include "wvd_timed/" isolates the timed region
        ms = timed(step_resident, args.steps)
        torch.cuda.nvtx.range_pop()
        launches = ops.LAUNCHES - launches0
        prof = ops.PROFILE
        ops.PROFILE = None
        clocks = sampler.stop() if rank == 0 else None
        torch.cuda.synchronize()
        att = [e0.elapsed_time(e1) for (e0, e1, _h, _n) in prof.get("self_attention", [])]
        att_heads = prof["self_attention"][0][2] if att else heads
        for _ in range(1):
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
        # the default CFG denoising step (posi + nega velocity, CFG combine, Euler update: wan_video_new.py:526-540) both
        # ways: the plain loop as the reference runs it, and with the loop-level fusion of SURVEY section 8(f)1 (text
        # embedding + cross-attention K/V cached across steps and branches, CFG + Euler as one device kernel, each velocity
        # prediction replayed as a CUDA graph).  Same outputs (tests/test_gpu_model.py); FLOPs per step still 2 x the full count.
        loop = None
        if not args.no_loop:
            nega = torch.zeros_like(devin["context"])
            lat0 = devin["latents"].clone()

            def cfg_steps(n, **kw):
                return V.denoise(dit, vace, lat0, devin["context"], nega, vace_context=devin.get("vace_context"), vace_scale=1.0,
                                 num_inference_steps=n, cfg_scale=5.0, use_unified_sequence_parallel=usp, **kw)
            loop = {}
            for name, kw in (("plain", dict(cache_text=False, fused_step=False, use_cuda_graph=False)),
                             ("fused", dict(cache_text=True, fused_step=True, use_cuda_graph=not usp))):
                n_steps = 3
                if name == "fused":
                    cache = V.TextCache()
                    gfn = V.GraphedModelFn(dit=dit, vace=vace, vace_scale=1.0, use_unified_sequence_parallel=usp, text_cache=cache) if not usp else None
                    sch = V.FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True)
                    sch.set_timesteps(50, shift=5.0)

                    def run_fused(n):
                        lat = lat0
                        for i in range(n):
                            ts_i = sch.timesteps[i].unsqueeze(0).to(dtype=torch.bfloat16, device=dev)
                            if gfn is not None:
                                vp = gfn(lat, ts_i, devin["context"], devin.get("vace_context")).clone()
                                vn = gfn(lat, ts_i, nega, devin.get("vace_context"))
                            else:
                                vp = V.model_fn_wan_video(dit=dit, vace=vace, latents=lat, timestep=ts_i, context=devin["context"], vace_context=devin.get("vace_context"), vace_scale=1.0, use_unified_sequence_parallel=usp, text_cache=cache)
                                vn = V.model_fn_wan_video(dit=dit, vace=vace, latents=lat, timestep=ts_i, context=nega, vace_context=devin.get("vace_context"), vace_scale=1.0, use_unified_sequence_parallel=usp, text_cache=cache)
                            lat = ops.cfg_euler_step(lat, vp, vn, 5.0, sch.dsigma(sch.timesteps[i]))
                        return lat
                    run_fused(1)                                    # capture / fill the caches outside the timed region
                    loop[name + "_s_per_cfg_step"] = timed(lambda: run_fused(n_steps), 1) / 1e3 / n_steps
                else:
                    cfg_steps(1, **kw)
                    loop[name + "_s_per_cfg_step"] = timed(lambda: cfg_steps(n_steps, **kw), 1) / 1e3 / n_steps
            loop["speedup"] = loop["plain_s_per_cfg_step"] / loop["fused_s_per_cfg_step"]
            loop["what"] = ("one default CFG denoising step = 2 velocity predictions + CFG combine + Euler update, 3 steps timed; "
                            "plain = the reference's loop on the wvd kernels; fused = text embedding + cross-attention K/V cached per "
                            "prompt, CFG + Euler in one kernel" + ("" if usp else ", CUDA-graph replay of each prediction") +
                            "; bit-identical outputs")
        gpu_ref = gpu_c1 = None
        if world == 1 and not args.no_gpu_reference:
            gpu_ref = gpu_reference_leg(dit, vace, devin, t_dev, step_resident(), ms, wl["size"])
        if world == 1 and not args.no_cpu_baseline and args.workload != "c1":
            # the wvd path at config c1 (the CPU leg's configuration), bf16 and fp32 parity mode, for a like-for-like ratio
            del dit, vace
            torch.cuda.empty_cache()
            gpu_c1 = {}
            for dt, nm in ((torch.bfloat16, "bf16"), (torch.float32, "f32")):
                d1, _ = S.build_models("1.3B", False, dev, dt, seed=0, lora_rank=None)
                i1 = {k: v.to(dev) for k, v in S.make_inputs(S.WORKLOADS["c1"]["latent"], with_vace=False, seed=1, dtype=dt, pin=False).items()}
                t1 = torch.tensor([1000.0], dtype=dt, device=dev)
                f1 = lambda: V.model_fn_wan_video(dit=d1, timestep=t1, **i1)     # noqa: E731
                for _ in range(3):
                    f1()
                gpu_c1[nm + "_s"] = timed(f1, 10) / 1e3
                del d1

    aux = None
    if world == 1 and not args.no_aux:
        aux = aux_leg(dev, peaks()["hbm"])
    if rank == 0:
        pk = peaks()
        att_ms = sum(att) / max(1, len(att))
        att_flops = S.attention_flops(tokens, att_heads)
        achieved = att_flops / (att_ms * 1e-3) / 1e12 if att else None
        h2d = sum(v.numel() * v.element_size() for v in host.values()) + t_host.numel() * 2
        line = dict(metric=METRIC, value=ms / 1e3, unit="s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms, higher_is_better=False, scaling="strong", vs_baseline=None, dtype="bf16",
                    data="synthetic", impl="wvd",
                    config=dict(workload=workload_name(args.workload, tokens), tokens=tokens,
                                parallelism=(f"ulysses{world} ({exchange_kind()})" if world > 1 else "single"),
                                l2="inputs larger than L2: the 34.6 GB of weights are streamed from HBM every step",
                                step="one model_fn_wan_video call; the default CFG denoising step is two"),
                    tokens_per_s=tokens / (ms / 1e3),
                    model_tflops_per_s=full_flops / (ms / 1e3) / 1e12 / 1.0,
                    tc_frac_of_measured_burst=full_flops / (ms / 1e3) / 1e12 / world / pk["bf16_burst"],
                    e2e=dict(value=ms_e2e / 1e3, unit="s", h2d_bytes_per_step=h2d,
                             d2h_bytes_per_step=out_host.numel() * 2),
                    gpu_launches=launches,
                    clocks=clocks,
                    roofline=dict(bound="tensor", kernel="wvd::attn4::attention_cg2p_kernel (self-attention, persistent cta_group::2)",
                                  achieved=achieved, peak=pk["bf16_sustained"], unit="TFLOP/s",
                                  frac=(achieved / pk["bf16_sustained"]) if achieved else None,
                                  traffic=ncu_traffic_bytes() if args.workload == "c3" and world == 1 else None,
                                  traffic_unit=f"bytes per launch (dram read + write, ncu --set full, profiles/{NCU_ATTENTION_SUMMARY}); "
                                               "algorithmic minimum 4*A = 1.214e9 (q, k, v read + out written once)",
                                  peak_source=f"{pk['src']} bf16_tflops_sustained (kernel timed inside a long step)",
                                  launches_timed=len(att), avg_launch_ms=att_ms,
                                  flops_per_launch=att_flops,
                                  share_of_step=(sum(att) / args.steps) / ms if att else None))
        if args.layers is not None:
            line["invalid"] = f"debug run with {args.layers} layers"
        if parity is not None:
            line["parity"] = parity
        if loop is not None:
            line["denoise_loop"] = loop
        if gpu_ref is not None:
            line["gpu_reference"] = gpu_ref
        if aux is not None:
            line["aux"] = aux
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_entry(S.model_flops("14B", 29640, True))
            if gpu_c1 is not None:
                gpu_c1["what"] = "this path at the CPU leg's configuration (c1: 1.3B, 1,280 tokens), resident inputs, 10 calls"
                gpu_c1["speedup_vs_cpu_c1_f32"] = line["cpu_baseline"]["value"] / gpu_c1["f32_s"]
                line["gpu_c1"] = gpu_c1
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
