"""Ulysses sequence parallelism over torch.distributed (NCCL on NVLink5/NVSwitch; gloo in the CPU tests).

Replaces ``diffsynth/distributed/xdit_context_parallel.py`` (xfuser ``xFuserLongContextAttention`` with
ulysses_degree = world, ring_degree = 1, wan_video_new.py:318-322) with an in-tree exchange:

  tokens are sharded outside attention (rank r owns rows [r*n_loc, (r+1)*n_loc) of x, of the VACE stream and of
  every hint -- unlike the reference, the VACE branch is sharded too); inside self-attention the
  heads are sharded: one all-to-all moves q|k|v from (n_loc, 3, H, 128) to (N, 3, H/P, 128), attention runs over
  H/P heads x all N tokens, and the inverse all-to-all returns (n_loc, H*128) for the o-projection.

The pack kernel writes the dest-rank-major send buffer; the receive buffer is consumed in place by the attention
kernel (strided TMA views), and the attention output is already the return-trip send layout.
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
