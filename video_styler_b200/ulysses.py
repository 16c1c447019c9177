"""Ulysses sequence parallelism over torch.distributed (NCCL on NVLink5/NVSwitch; gloo in the CPU tests).

Replaces ``diffsynth/distributed/xdit_context_parallel.py`` (xfuser ``xFuserLongContextAttention`` with
ulysses_degree = world, ring_degree = 1, wan_video_new.py:318-322) with an in-tree exchange:

  tokens are sharded outside attention (rank r owns rows [r*n_loc, (r+1)*n_loc) of x, of the VACE stream and of
  every hint -- unlike the reference, the VACE branch is sharded too); inside self-attention the
  heads are sharded: one all-to-all moves q|k|v from (n_loc, 3, H, 128) to (N, 3, H/P, 128), attention runs over
  H/P heads x all N tokens, and the inverse all-to-all returns (n_loc, H*128) for the o-projection.

Two implementations of the exchange:

  P2PUlyssesExchange   (default on NVLink boxes) both all-to-alls are FUSED into the kernels on either side of them
                       and run over peer memory: the pack kernel stores each destination's q|k|v chunk straight into
                       that rank's receive buffer (wvd_ulysses_scatter_qkv), and the attention kernel's epilogue stores
                       every query row straight into its owner's o-projection input (wvd_attention_fwd_scatter).  The
                       buffers come from a torch symmetric-memory rendezvous; two cross-rank barriers per attention
                       order the steps.  No send/receive staging, no unpack pass, no NCCL launch on the block path.
  UlyssesExchange      the NCCL baseline (``all_to_all_single``; gloo in the CPU tests): pack kernel -> all-to-all ->
                       attention on the receive buffer in place -> all-to-all -> unpack kernel.
Full-width QK-RMSNorm + RoPE run BEFORE the scatter with the rank's global token offset, as the reference does
(xdit_context_parallel.py:27-40, 110-117).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .engine import SelfAttnExchange


def shard_bounds(n_tokens: int, world: int, rank: int):
    """torch.chunk semantics of the reference (wan_video_new.py:1412-1417): shards of ceil(N/P) rows, the last one
    zero-padded.  Returns (lo, hi, n_loc): this rank owns global rows [lo, hi) stored in an n_loc-row buffer."""
    n_loc = -(-n_tokens // world)
    lo = min(rank * n_loc, n_tokens)
    hi = min(lo + n_loc, n_tokens)
    return lo, hi, n_loc


class UlyssesExchange(SelfAttnExchange):
    def __init__(self, group, n_tokens: int):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_tokens = n_tokens

    def attend(self, ops, qkv, heads: int, out, ws):
        p = self.world
        if heads % p != 0:
            raise ValueError(f"Ulysses needs num_heads ({heads}) divisible by the sequence-parallel world size ({p})")
        n_loc = qkv.shape[0]
        hl = heads // p
        w = hl * 128
        send = ops.ulysses_pack_qkv(qkv, heads, p, out=ws.get("a2a_send", (p, n_loc, 3, hl, 128)))
        recv = ws.get("a2a_recv", (p, n_loc, 3, hl, 128))
        dist.all_to_all_single(recv, send, group=self.group)
        r = recv.view(p * n_loc, 3 * w)
        n = self.n_tokens                       # rows >= n are the zero padding of the last shard: never attended
        o_heads = ws.get("a2a_out", (p * n_loc, w))
        ops.attention(r[:n, :w], r[:n, w:2 * w], r[:n, 2 * w:], hl, out=o_heads[:n])
        if p * n_loc > n:
            o_heads[n:].zero_()
        back = ws.get("a2a_back", (p, n_loc, w))
        dist.all_to_all_single(back, o_heads.view(p, n_loc, w), group=self.group)
        return ops.ulysses_unpack_out(back, heads, p, out=out)

    def all_gather_tokens(self, y_loc):
        """(n_loc, C) per rank -> (P*n_loc, C) in global token order (wan_video_new.py:1461)."""
        out = torch.empty((self.world * y_loc.shape[0], y_loc.shape[1]), dtype=y_loc.dtype, device=y_loc.device)
        dist.all_gather_into_tensor(out, y_loc.contiguous(), group=self.group)
        return out


class P2PUlyssesExchange(UlyssesExchange):
    """Ulysses exchange with both all-to-alls fused into the producing kernels over NVLink peer memory (see the module
    docstring).  Bit-identical to UlyssesExchange (same arithmetic, only the data path differs)."""

    def __init__(self, group, n_tokens: int):
        super().__init__(group, n_tokens)
        self._bufs = {}
        self._nccl_only = False
        self._probed = None

    def _buffers(self, n_loc: int, heads: int, device):
        key = (n_loc, heads)
        if key not in self._bufs:
            import importlib
            symm_mem = importlib.import_module("torch.distributed._symmetric_memory")
            p, w = self.world, (heads // self.world) * 128
            recv = symm_mem.empty((p * n_loc, 3 * w), dtype=torch.bfloat16, device=device)
            aout = symm_mem.empty((n_loc, heads * 128), dtype=torch.bfloat16, device=device)
            recv.zero_()
            aout.zero_()          # rows past the last token (ragged N) are never written: they stay zero
            grp = self.group if self.group is not None else dist.group.WORLD
            h_recv = symm_mem.rendezvous(recv, grp)
            h_out = symm_mem.rendezvous(aout, grp)
            torch.cuda.synchronize(device)
            h_recv.barrier(channel=0)
            self._bufs[key] = (recv, aout, h_recv, h_out, [int(x) for x in h_recv.buffer_ptrs], [int(x) for x in h_out.buffer_ptrs])
        return self._bufs[key]

    _v_sent = False

    def v_ready(self, ops, qkv, heads: int) -> None:
        """The v columns of ``qkv`` are final on the current stream: store them into the peers on the side stream.
        Safe to overwrite the peers' receive buffers: every rank passed the previous attention's second barrier only
        after its attention had finished reading them.  No-op (the NCCL path packs everything later) when the
        peer-memory buffers are unavailable."""
        if heads % self.world != 0 or qkv.dtype != torch.bfloat16:
            return
        bufs = self._buffers_or_none(qkv.shape[0], heads, qkv.device)
        if bufs is None:
            return
        main = torch.cuda.current_stream()
        side = self._side_stream(qkv.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ops.ulysses_scatter_v(qkv, heads, bufs[4], self.rank)
        self._v_sent = True

    def _side_stream(self, device):
        st = getattr(self, "_side", None)
        if st is None:
            st = self._side = torch.cuda.Stream(device=device)
        return st

    def _probe(self, device) -> bool:
        """Can EVERY rank take the peer-memory path?  Local, failure-free checks (symmetric-memory module importable,
        peer access to every GPU of the group) agreed on with one all-reduce BEFORE anybody enters the rendezvous:
        the rendezvous is itself a collective, so a rank that failed on its own on the way in would leave the others
        blocked inside it.  After a unanimous yes, a failure inside the rendezvous is not recoverable by falling back
        (the peers may already be waiting in it) and is raised."""
        ok = 1
        try:
            import importlib
            importlib.import_module("torch.distributed._symmetric_memory")
            me = torch.device(device).index
            me = torch.cuda.current_device() if me is None else me
            devs = [None] * self.world
            dist.all_gather_object(devs, (me, _host_id()), group=self.group)
            for d, host in devs:
                if host != _host_id() or (d != me and not torch.cuda.can_device_access_peer(me, d)):
                    ok = 0
        except Exception as e:      # noqa: BLE001 -- any local failure of the optional fast path selects the baseline
            ok = 0
            import warnings
            warnings.warn(f"peer-memory Ulysses exchange unavailable on this rank ({e!r})")
        flag = torch.tensor([ok], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        return int(flag.item()) == 1

    def _buffers_or_none(self, n_loc: int, heads: int, device):
        """The symmetric buffers, or None if this box cannot take the peer-memory path (agreed on by all ranks in
        ``_probe``); the exchange then stays on NCCL collectives for good -- still the GPU path, just not the fused one."""
        if self._nccl_only:
            return None
        key = (n_loc, heads)
        if key in self._bufs:
            return self._bufs[key]
        if self._probed is None:
            self._probed = self._probe(device)
        if not self._probed:
            self._nccl_only = True
            import warnings
            warnings.warn("peer-memory Ulysses exchange unavailable on some rank; using NCCL all_to_all")
            return None
        return self._buffers(n_loc, heads, device)        # a failure in here aborts the job (see _probe)

    def norm_rope_attend(self, ops, qkv, heads: int, wq, wk, eps: float, rope, out, ws):
        """The q / k part of the first all-to-all fused into the RoPE kernel's stores, the v part as a one-third
        scatter: no pass over q and k between the RoPE kernel and the attention."""
        p = self.world
        if heads % p != 0:
            raise ValueError(f"Ulysses needs num_heads ({heads}) divisible by the sequence-parallel world size ({p})")
        per_token = tuple(rope.grid) == (0, 0, 0)
        bufs = None if (per_token or qkv.dtype != torch.bfloat16) else self._buffers_or_none(qkv.shape[0], heads, qkv.device)
        if bufs is None:
            if self._v_sent:            # announced, but this call takes the collective path after all: just rejoin
                torch.cuda.current_stream().wait_stream(self._side_stream(qkv.device))
                self._v_sent = False
            return super().norm_rope_attend(ops, qkv, heads, wq, wk, eps, rope, out, ws)
        d = heads * 128
        recv_ptrs = bufs[4]
        # Safe to overwrite the peers' receive buffers: every rank passed the previous attention's second barrier only
        # after its attention had finished reading them.
        # The v third went out on the side stream right after its projection (v_ready), under the q | k projection; if
        # the caller did not announce it, it goes now, still concurrently with the RoPE kernel's stores.
        main = torch.cuda.current_stream()
        side = self._side_stream(qkv.device)
        if not self._v_sent:
            self.v_ready(ops, qkv, heads)
        self._v_sent = False
        ops.qk_rmsnorm_rope_scatter(qkv[:, :d], qkv[:, d:2 * d], wq, wk, eps, rope.table, rope.grid, rope.token_offset,
                                    rope.frame_ids, recv_ptrs, self.rank)
        main.wait_stream(side)
        return self._attend_received(ops, bufs, heads, qkv.shape[0])

    def attend(self, ops, qkv, heads: int, out, ws):
        p, r = self.world, self.rank
        if heads % p != 0:
            raise ValueError(f"Ulysses needs num_heads ({heads}) divisible by the sequence-parallel world size ({p})")
        n_loc = qkv.shape[0]
        bufs = self._buffers_or_none(n_loc, heads, qkv.device)
        if bufs is None:
            return super().attend(ops, qkv, heads, out, ws)
        # (1) q|k|v -> every rank's receive buffer.  Safe to overwrite: every rank passed the previous barrier (2)
        #     only after its previous attention had finished reading.
        ops.ulysses_scatter_qkv(qkv, heads, bufs[4], r)
        return self._attend_received(ops, bufs, heads, n_loc)

    def _attend_received(self, ops, bufs, heads: int, n_loc: int):
        p, r = self.world, self.rank
        hl = heads // p
        w = hl * 128
        recv, aout, h_recv, h_out, recv_ptrs, out_ptrs = bufs
        _device_barrier(h_recv, "after scatter")
        n = self.n_tokens                       # rows >= n are the zero padding of the last shard: never attended
        # (2) attention over my heads and all tokens; rows go straight to their owners' o-projection input.  Safe to
        #     overwrite: every rank entered barrier (1) only after its previous o-projection had been enqueued
        #     ahead of it on its stream.
        ops.attention_scatter(recv[:n, :w], recv[:n, w:2 * w], recv[:n, 2 * w:], hl, out_ptrs, heads * 128, n_loc, r * w)
        _device_barrier(h_out, "after attention")
        return aout


def _device_barrier(handle, what: str) -> None:
    """Cross-rank barrier on the current stream (symmetric-memory signal pads: a kernel, no host synchronisation).
    One call site per barrier kind so that tools/step_breakdown.py can bracket them."""
    handle.barrier(channel=0)


def _host_id() -> str:
    import socket
    return socket.gethostname()


_EXCHANGES = {}


def make_exchange(group, n_tokens: int, device) -> UlyssesExchange:
    """The exchange for this (group, token count): peer-memory fused kernels when the group runs on CUDA with
    symmetric memory available (WVD_ULYSSES=nccl forces the NCCL baseline), else NCCL / gloo collectives.  Cached:
    the symmetric-memory rendezvous is a collective and is done once."""
    import os
    key = (id(group), n_tokens, str(device))
    if key in _EXCHANGES:
        return _EXCHANGES[key]
    want = os.environ.get("WVD_ULYSSES", "p2p").lower()
    ex = None
    if want != "nccl" and torch.device(device).type == "cuda" and dist.get_backend(group) == "nccl":
        try:
            import importlib
            importlib.import_module("torch.distributed._symmetric_memory")
            ex = P2PUlyssesExchange(group, n_tokens)
        except Exception:           # no symmetric memory in this torch build: NCCL collectives
            ex = None
    if ex is None:
        ex = UlyssesExchange(group, n_tokens)
    _EXCHANGES[key] = ex
    return ex
