"""Multi-GPU parity: P-rank Ulysses output vs the single-GPU output of the same model on the same inputs.

    torchrun --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29511 tools/ulysses_parity.py \
        [--cases c3w,c5w,ragged,heads12] [--layers 2] [--json gpurun_out/ulysses_parity.json]

Ulysses is mathematically exact (same heads, same softmax, the same kernels on the same operands), so the sharded output
is expected to be BIT-IDENTICAL to the unsharded one; the assertion is the single-GPU bf16 tolerance of BASELINE.json
(cos >= 0.999, relL2 <= 1e-2) and the bit-identity is reported.  tests/test_gpu_multi.py spawns this over all visible
GPUs.  Reference: diffsynth/distributed/xdit_context_parallel.py:110-131, wan_video_new.py:1412-1417,1447-1449,1459-1462.

Cases (all random-init, reduced depth so that the run takes seconds):
  c3w      14B width + VACE + LoRA stand-in at the c3 grid (29,640 tokens: 14,820 / 7,410 / 3,705 per rank)
  c5w      14B width + VACE at the c5 grid (75,600 tokens), attention-bound
  ragged   14B width + VACE on a grid whose token count (3 x 15 x 13 = 585) does not divide by P: the last shard is
           zero-padded (wan_video_new.py:1414-1416), the padding is never attended
  heads12  the 1.3B model (12 heads) at P = 8 must raise on every rank (heads do not divide), not hang
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import video_styler_b200 as V  # noqa: E402
from video_styler_b200 import synthetic as S  # noqa: E402
from video_styler_b200 import ulysses as U  # noqa: E402

CASES = {
    "c3w": dict(size="14B", vace=True, latent=(1, 16, 19, 60, 104)),
    "c5w": dict(size="14B", vace=True, latent=(1, 16, 21, 90, 160)),
    "ragged": dict(size="14B", vace=True, latent=(1, 16, 3, 30, 26)),
    "heads12": dict(size="1.3B", vace=False, latent=(1, 16, 3, 16, 16)),
}


def run_case(name, layers, dev, rank, world):
    c = CASES[name]
    dit, vace = S.build_models(c["size"], c["vace"], dev, torch.bfloat16, seed=0, lora_rank=128, num_layers=layers)
    inp = {k: v.to(dev) for k, v in S.make_inputs(c["latent"], with_vace=c["vace"], seed=1, pin=False).items()}
    ts = torch.tensor([832.0], dtype=torch.bfloat16, device=dev)
    heads = S.DIT_CONFIGS[c["size"]]["num_heads"]
    b, ch, f, h, w = c["latent"]
    tokens = f * (h // 2) * (w // 2)
    res = dict(case=name, tokens=tokens, heads=heads, world=world, layers=layers)
    with torch.no_grad():
        if heads % world != 0:
            try:
                V.model_fn_wan_video(dit=dit, vace=vace, timestep=ts, vace_scale=1.0, use_unified_sequence_parallel=True, **inp)
                res.update(ok=False, error="no exception for heads % world != 0")
            except ValueError as e:
                res.update(ok=True, raised=str(e))
            torch.cuda.synchronize()
            return res
        single = V.model_fn_wan_video(dit=dit, vace=vace, timestep=ts, vace_scale=1.0, **inp)
        sharded = V.model_fn_wan_video(dit=dit, vace=vace, timestep=ts, vace_scale=1.0, use_unified_sequence_parallel=True, **inp)
    torch.cuda.synchronize()
    d = sharded.double() - single.double()
    rel = float(d.norm() / single.double().norm())
    cos = float(torch.nn.functional.cosine_similarity(sharded.double().flatten(), single.double().flatten(), dim=0))
    ref = sharded.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.equal(ref, sharded))                  # every rank holds the same gathered result
    kind = ",".join(sorted({("UlyssesExchange" if getattr(e, "_nccl_only", False) else type(e).__name__)
                            for e in U._EXCHANGES.values()}))
    res.update(rel_l2=rel, cos=cos, max_abs=float(d.abs().max()), identical_across_ranks=same,
               bit_identical=bool(torch.equal(sharded, single)), exchange=kind, ragged=tokens % world != 0,
               ok=bool(rel <= 1e-2 and cos >= 0.999 and same and torch.isfinite(sharded.float()).all()))
    del dit, vace
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--cases", default="c3w,ragged")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    results = []
    for name in a.cases.split(","):
        r = run_case(name, a.layers, dev, rank, world)
        oks = [None] * world
        dist.all_gather_object(oks, bool(r["ok"]))
        r["ok_all_ranks"] = all(oks)
        results.append(r)
        if rank == 0:
            print(json.dumps(r), flush=True)
    ok = all(r["ok_all_ranks"] for r in results)
    if rank == 0 and a.json:
        os.makedirs(os.path.dirname(os.path.abspath(a.json)), exist_ok=True)
        with open(a.json, "w") as fh:
            json.dump(dict(world=world, ok=ok, results=results), fh, indent=1)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
