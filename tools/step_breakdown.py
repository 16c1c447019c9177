"""Where one model_fn_wan_video step goes, per kernel family, at N ranks (CUDA events on the launching stream).

    python tools/step_breakdown.py [--workload c3] [--steps 3] [--json out.json]
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/step_breakdown.py ...

Every ops.* entry point (one C-ABI kernel each) and every symmetric-memory barrier of the Ulysses exchange is bracketed
by a CUDA-event pair while K steps run; per family the tool prints launches / step, ms / step (sum of the launch
durations) and the share of the step, then "gaps" = step time - sum of the bracketed durations: launch latency, the
PyTorch glue kernels (copies, embeddings) and idle time between kernels.  With events between all kernels the step itself
runs a little slower than in bench.py (the untouched step time is printed beside it).  Rank 0 prints; the table holds the
MAX over ranks per family (a barrier's duration on one rank is the time it waited for the slowest peer).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import video_styler_b200 as V  # noqa: E402
from video_styler_b200 import ops, synthetic as S  # noqa: E402
from video_styler_b200 import ulysses as U  # noqa: E402

OPS = ["ln_modulate", "qk_rmsnorm_rope", "linear", "linear_grouped", "attention", "scale_add", "ulysses_scatter_qkv",
       "ulysses_scatter_v", "qk_rmsnorm_rope_scatter", "attention_scatter", "ulysses_pack_qkv", "ulysses_unpack_out"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--graph", action="store_true", help="also time the step replayed as a CUDA graph")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    wl = S.WORKLOADS[a.workload]
    dit, vace = S.build_models(wl["size"], wl["vace"], dev, torch.bfloat16, seed=0, lora_rank=128, num_layers=a.layers)
    inp = {k: v.to(dev) for k, v in S.make_inputs(wl["latent"], with_vace=wl["vace"], seed=1, pin=False).items()}
    ts = torch.tensor([832.0], dtype=torch.bfloat16, device=dev)
    usp = world > 1

    def step():
        return V.model_fn_wan_video(dit=dit, vace=vace, timestep=ts, vace_scale=1.0, use_unified_sequence_parallel=usp, **inp)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        sync()
        t = torch.tensor([e0.elapsed_time(e1) / k], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    with torch.no_grad():
        for _ in range(3):
            step()
        plain_ms = timed(step, a.steps)
        graph_ms = None
        # ---- bracket every op ----
        records = {}

        def wrap(mod, name, family=None):
            fn = getattr(mod, name)

            def inner(*args, **kw):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = fn(*args, **kw)
                e1.record()
                fam = family(args, kw) if callable(family) else (family or name)
                records.setdefault(fam, []).append((e0, e1))
                return r
            setattr(mod, name, inner)
            return fn

        def linear_family(args, kw):
            x, w = args[0], args[1]
            return f"linear {w.shape[0]}x{w.shape[1]}" + (" (M=512 text)" if x.shape[0] <= 512 else "")

        def attn_family(args, kw):
            return "attention self" if args[0].shape[0] == args[1].shape[0] else "attention cross"

        def rope_family(args, kw):
            return "qk_rmsnorm_rope (self q,k)" if args[1] is not None else "qk_rmsnorm (cross q / k)"

        saved = []
        for name in OPS:
            fam = {"linear": linear_family, "attention": attn_family, "qk_rmsnorm_rope": rope_family}.get(name)
            saved.append((ops, name, wrap(ops, name, fam)))
        # the exchange's device-side barriers
        orig_barrier = U._device_barrier

        def barrier(handle, what):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            orig_barrier(handle, what)
            e1.record()
            records.setdefault("barrier " + what, []).append((e0, e1))
        U._device_barrier = barrier
        step()
        records.clear()
        inst_ms = timed(step, a.steps)
        for mod, name, fn in saved:
            setattr(mod, name, fn)
        U._device_barrier = orig_barrier

    graph_err = None
    if a.graph:          # last: a failed capture can leave the stream unusable
        try:
            with torch.no_grad():
                g = V.GraphedModelFn(dit=dit, vace=vace, vace_scale=1.0, use_unified_sequence_parallel=usp)
                gf = lambda: g(inp["latents"], ts, inp["context"], inp.get("vace_context"))     # noqa: E731
                gf()
                graph_ms = timed(gf, a.steps)
        except Exception as e:      # noqa: BLE001
            graph_err = repr(e)[:300]
    fams = sorted(records)
    tot = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in records[f]) / a.steps for f in fams], device=dev, dtype=torch.float64)
    cnt = [len(records[f]) / a.steps for f in fams]
    tot_max = tot.clone()
    if world > 1:
        dist.all_reduce(tot_max, op=dist.ReduceOp.MAX)
    if rank == 0:
        rows = sorted(zip(fams, cnt, tot.tolist(), tot_max.tolist()), key=lambda r: -r[3])
        covered = sum(r[2] for r in rows)
        print(f"# {a.workload}, {world} GPU(s): step {plain_ms:.2f} ms untouched" + (f", {graph_ms:.2f} ms as a CUDA graph" if graph_ms else "") +
              f", {inst_ms:.2f} ms with an event pair around every op")
        print(f"# {'family':42s} {'launches/step':>13s} {'ms/step rank0':>14s} {'max over ranks':>15s} {'share of step':>14s}")
        for f, c, t, tm in rows:
            print(f"  {f:42s} {c:13.1f} {t:14.3f} {tm:15.3f} {100 * t / inst_ms:13.1f}%")
        print(f"  {'(sum of the bracketed ops, rank 0)':42s} {sum(cnt):13.1f} {covered:14.3f} {'':15s} {100 * covered / inst_ms:13.1f}%")
        print(f"  {'gaps: glue kernels, launch latency, idle':42s} {'':13s} {inst_ms - covered:14.3f} {'':15s} {100 * (inst_ms - covered) / inst_ms:13.1f}%")
        if graph_err:
            print("# CUDA-graph capture failed:", graph_err)
        if a.json:
            json.dump(dict(workload=a.workload, world=world, plain_ms=plain_ms, graph_ms=graph_ms, instrumented_ms=inst_ms,
                           families=[dict(name=f, launches=c, ms=t, ms_max=tm) for f, c, t, tm in rows]), open(a.json, "w"), indent=1)
    sys.stdout.flush()
    torch.cuda.synchronize()
    if a.graph:
        os._exit(0)      # destroying an NCCL communicator that live CUDA graphs still reference hangs: just leave
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
