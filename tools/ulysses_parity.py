"""Multi-GPU parity: P-rank Ulysses output vs the single-GPU output of the same model on the same inputs.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ulysses_parity.py [--layers 6]
Ulysses is mathematically exact (same heads, same softmax), so the tolerance is the single-GPU bf16 tolerance.
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import video_styler_b200 as V  # noqa: E402
from video_styler_b200 import synthetic as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=6)
ap.add_argument("--workload", default="c3")
a = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
wl = S.WORKLOADS[a.workload]
dit, vace = S.build_models(wl["size"], wl["vace"], dev, torch.bfloat16, seed=0, lora_rank=128, num_layers=a.layers)
inp = {k: v.to(dev) for k, v in S.make_inputs(wl["latent"], with_vace=wl["vace"], seed=1, pin=False).items()}
ts = torch.tensor([832.0], dtype=torch.bfloat16, device=dev)
with torch.no_grad():
    single = V.model_fn_wan_video(dit=dit, vace=vace, timestep=ts, vace_scale=1.0, **inp)
    sharded = V.model_fn_wan_video(dit=dit, vace=vace, timestep=ts, vace_scale=1.0, use_unified_sequence_parallel=True, **inp)
torch.cuda.synchronize()
d = (sharded.double() - single.double())
rel = float(d.norm() / single.double().norm())
cos = float(torch.nn.functional.cosine_similarity(sharded.double().flatten(), single.double().flatten(), dim=0))
# all ranks must hold the same gathered result
ref = sharded.clone()
dist.broadcast(ref, 0)
same = bool(torch.equal(ref, sharded))
from video_styler_b200 import ulysses as _U  # noqa: E402
kind = ",".join(sorted({type(e).__name__ for e in _U._EXCHANGES.values()}))
print(f"rank {rank}/{world} [{kind}]: ulysses vs single-GPU rel_l2 {rel:.3e} cos {cos:.7f} max_abs {float(d.abs().max()):.3e} identical_across_ranks {same}", flush=True)
bit = bool(torch.equal(sharded, single))
print(f"rank {rank}: bit-identical to single-GPU: {bit}", flush=True)
ok = rel <= 1e-2 and cos >= 0.999 and same
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
