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
             the oracle's restatement of the reference DiTBlock on the host cores on a bounded sample (one 14B block,
             one latent frame = 1,560 tokens, fp32), scaled by algorithmic FLOPs to the full call
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


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE self-attention launch at c3, from the committed
    `ncu --set full` capture summary (profiles/, written by tools/ncu_summary.py).  None if the summary is absent."""
    p = os.path.join(ROOT, "profiles", "r1_attention_pair_c3.txt")
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
# CPU leg: the oracle (port of the reference path) on the host cores, bounded sample
# ----------------------------------------------------------------------------------------------------------------
def cpu_sample(steps=3, warmup=1):
    import torch
    from oracle import wan_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.DIT_CONFIGS["14B"]
    d, ffn, heads = cfg["dim"], cfg["ffn_dim"], cfg["num_heads"]
    n, lctx = 1560, 512                       # one latent frame of c3 (30 x 52 tokens)
    shapes = O._block_shapes("blocks.0.", d, ffn)
    sd = O.make_state_dict(shapes, seed=0)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, n, d, generator=g)
    ctx = torch.randn(1, lctx, d, generator=g)
    t_mod = torch.randn(1, 6, d, generator=g) * 0.1
    freqs = O.rope_freqs(128, 1, 30, 52)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.dit_block(sd, "blocks.0.", x, ctx, t_mod, freqs, heads, cfg["eps"])
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    nn, l = float(n), float(lctx)
    sample_flops = 8 * nn * d * d + 4 * nn * nn * d + 4 * nn * d * d + 4 * l * d * d + 4 * nn * l * d + 4 * nn * d * ffn
    return statistics.median(times), sample_flops, torch.get_num_threads()


def cpu_baseline_entry(full_flops, steps=3, warmup=1):
    t, fl, cores = cpu_sample(steps, warmup)
    return dict(value=t * full_flops / fl, unit="s", cores=cores, kind="port",
                sample=f"oracle DiTBlock (14B width, fp32) on 1,560 tokens = one latent frame: {t:.2f} s median for "
                       f"{fl/1e12:.3f} TFLOP, scaled by algorithmic FLOPs ({full_flops/1e12:.1f} TFLOP per call; the "
                       f"N^2 attention term makes this a lower bound)",
                sample_seconds=t, sample_tflops_per_s=fl / t / 1e12)


def run_reference(args, rank):
    if rank != 0:
        return
    from video_styler_b200 import synthetic as S
    wl = S.WORKLOADS[args.workload]
    b, c, f, h, w = wl["latent"]
    tokens = f * (h // 2) * (w // 2)
    full = S.model_flops(wl["size"], tokens, wl["vace"])
    t0 = time.perf_counter()
    entry = cpu_baseline_entry(full, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    line = dict(metric=METRIC, value=entry["value"], unit="s", n_gpus=0, steps=args.steps, warmup=args.warmup,
                ms_per_step=entry["value"] * 1e3, higher_is_better=False, scaling="strong", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=workload_name(args.workload, tokens), tokens=tokens),
                cpu_baseline=entry,
                e2e=dict(value=entry["value"], unit="s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                wall_s=time.perf_counter() - t0)
    print(json.dumps(line), flush=True)


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

    with torch.no_grad():
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
                    roofline=dict(bound="tensor", kernel="wvd::attn2::attention_pair_kernel (self-attention)",
                                  achieved=achieved, peak=pk["bf16_sustained"], unit="TFLOP/s",
                                  frac=(achieved / pk["bf16_sustained"]) if achieved else None,
                                  traffic=ncu_traffic_bytes() if args.workload == "c3" and world == 1 else None,
                                  traffic_unit="bytes per launch (dram read + write, ncu --set full, profiles/r1_attention_pair_c3.txt); "
                                               "algorithmic minimum 4*A = 1.214e9 (q, k, v read + out written once)",
                                  peak_source=f"{pk['src']} bf16_tflops_sustained (kernel timed inside a long step)",
                                  launches_timed=len(att), avg_launch_ms=att_ms,
                                  flops_per_launch=att_flops,
                                  share_of_step=(sum(att) / args.steps) / ms if att else None))
        if args.layers is not None:
            line["invalid"] = f"debug run with {args.layers} layers"
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_entry(S.model_flops(wl["size"], tokens, wl["vace"]))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
