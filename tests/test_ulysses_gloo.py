"""Ulysses host logic at world_size 2 over gloo on CPU: the sharded path (tokens sharded outside attention, heads
inside; VACE sharded too) must reproduce the single-process golden output of the REAL reference."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, latent_shape, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import wan_oracle as O
        from tests import cpu_backend
        from tests.test_host_logic_cpu import build_models, run_model_fn
        if name is not None:
            fix = torch.load(os.path.join(ROOT, "tests", "golden", name + ".pt"), weights_only=False)
            ref = fix["output"]
        else:
            fix = dict(size="tiny", seeds=dict(dit=0, vace=3, lora=2, inputs=1), perturb=True, weight_scale=1.0,
                       with_vace=True, lora=False, latent_shape=latent_shape, timestep=640.0)
            cfg, vcfg = O.DIT_CONFIGS["tiny"], O.VACE_CONFIGS["tiny"]
            sd = O.make_state_dict(O.dit_param_shapes(cfg), seed=0, perturb_norms=True)
            vsd = O.make_state_dict(O.vace_param_shapes(vcfg), seed=3, perturb_norms=True)
            inp = O.make_inputs(latent_shape, cfg["text_dim"], seed=1, with_vace=True)
            with torch.no_grad():
                ref = O.model_fn_wan_video(sd, cfg, inp["latents"], torch.tensor([640.0]), inp["context"], vsd, vcfg,
                                           inp["vace_context"], 1.0)
        dit, vace = build_models(fix)
        out = run_model_fn(fix, dit, vace, cpu_backend, use_unified_sequence_parallel=True)
        m = O.parity_metrics(out, ref)
        q.put((rank, m))
    finally:
        dist.destroy_process_group()


def _run(world, name, latent_shape=None):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, latent_shape, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    return dict(res)


@pytest.mark.parametrize("name", ["tiny_vace_lora", "small_vace"])
def test_ulysses_world2_matches_single_process_reference(name):
    for rank, m in _run(2, name).items():
        assert m["max_abs"] <= 5e-5 and m["rel_l2"] <= 2e-5, (rank, m)


def test_ulysses_ragged_tokens_world2():
    """45 tokens over 2 ranks: the last shard is zero-padded (wan_video_new.py:1414-1416) and the padding is never
    attended, so the result still equals the unsharded oracle."""
    for rank, m in _run(2, None, (1, 16, 3, 6, 10)).items():
        assert m["max_abs"] <= 5e-5 and m["rel_l2"] <= 2e-5, (rank, m)
