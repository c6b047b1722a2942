"""Multi-GPU plumbing: one process per GPU, NCCL through torch.distributed.

The path shards by independent 30 s chunks (chunk i -> rank i mod W), so there is no collective on the data path:
  * start-up: rank 0 holds the packed weights, everyone else receives the two arenas with one broadcast each
    (3.09 GB bf16 + a few MB f32 for large-v3);
  * end: fixed-size per-chunk results {tokens i32[448], token_ts f32[448], len i32} are all-gathered.
The same code runs on `gloo` with CPU tensors for the world_size-2 tests (tests/test_distributed_cpu.py)."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import weights as Wt

N_TOK = 448


def shard_round_robin(n_items: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_items, world))


def broadcast_weights(packed: Optional[Wt.PackedWeights], config: Dict, device, src: int = 0) -> Wt.PackedWeights:
    """Every rank returns PackedWeights on `device`; only rank `src` needs to pass them in."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert packed is not None
        return packed
    if dist.get_rank() == src:
        assert packed is not None
        pw = packed if packed.arena_bf16.device == torch.device(device) else packed.to(device)
    else:
        pw = Wt.empty_packed(config, device)
    dist.broadcast(pw.arena_bf16, src=src)
    dist.broadcast(pw.arena_f32, src=src)
    return pw


def pack_result(tokens: np.ndarray, token_ts: np.ndarray, n_tok: int = N_TOK) -> Tuple[np.ndarray, np.ndarray]:
    """-> (int32 [n_tok + 1] = tokens padded with -1 then the length, float32 [n_tok]).  A chunk that does not fit the
    record is an error (a chunk that needed several seek passes can exceed 448 tokens: gather_results sizes the record from
    the all-reduced maximum, so nothing is ever truncated silently)."""
    n = len(tokens)
    if n > n_tok:
        raise ValueError(f"pack_result: {n} tokens do not fit a record of {n_tok}")
    t = np.full(n_tok + 1, -1, np.int32)
    t[:n] = tokens[:n]
    t[n_tok] = n
    ts = np.zeros(n_tok, np.float32)
    m = min(n, len(token_ts))
    ts[:m] = token_ts[:m]
    return t, ts


def gather_results(local: List[Tuple[np.ndarray, np.ndarray]], n_items: int, device) -> Optional[List[Tuple[np.ndarray, np.ndarray]]]:
    """local: results of this rank's round-robin shard, in shard order.  Returns the full list in item order on
    every rank (all_gather of fixed-size records; ~3.6 KB per chunk; the record length is the all-reduced maximum
    token count, at least 448)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    per_rank = (n_items + world - 1) // world
    n_tok = max([N_TOK] + [len(tok) for tok, _ in local])
    if world > 1:
        nt = torch.tensor([n_tok], dtype=torch.int64, device=device)
        dist.all_reduce(nt, op=dist.ReduceOp.MAX)
        n_tok = int(nt.item())
    ti = torch.full((per_rank, n_tok + 1), -1, dtype=torch.int32)
    tf = torch.zeros((per_rank, n_tok), dtype=torch.float32)
    for k, (tok, ts) in enumerate(local):
        a, b = pack_result(tok, ts, n_tok)
        ti[k] = torch.from_numpy(a)
        tf[k] = torch.from_numpy(b)
    ti, tf = ti.to(device), tf.to(device)
    if world == 1:
        gi, gf = [ti], [tf]
    else:
        gi = [torch.empty_like(ti) for _ in range(world)]
        gf = [torch.empty_like(tf) for _ in range(world)]
        dist.all_gather(gi, ti)
        dist.all_gather(gf, tf)
    out: List[Optional[Tuple[np.ndarray, np.ndarray]]] = [None] * n_items
    for r in range(world):
        ai, af = gi[r].cpu().numpy(), gf[r].cpu().numpy()
        for k, item in enumerate(shard_round_robin(n_items, r, world)):
            n = int(ai[k, n_tok])
            out[item] = (ai[k, :n].astype(np.int64), af[k, :n].copy())
    return out
