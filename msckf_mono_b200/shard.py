"""Sharding of independent filters / sequences over ranks.

One process per GPU; sequence i belongs to rank i mod world; no collective on the data path -- the only
communication is the barrier and the gathering of per-sequence summary metrics (a few hundred bytes), over NCCL on
the GPU box and gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def sequences_of_rank(n_sequences: int, rank: int, world: int):
    """round-robin: sequence i -> rank i mod world (64 sequences on 8 GPUs -> 8 each)."""
    return list(range(rank, n_sequences, world))


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(values, device="cpu"):
    """element-wise max of a list of floats over all ranks (the slowest rank defines the time)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def gather_summaries(local: dict):
    """per-sequence summaries {seq: {...}} of every rank, merged (rank order) on all ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(local)
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, local)
    merged = {}
    for part in out:
        for k, v in part.items():
            if k in merged:
                raise RuntimeError(f"sequence {k} was processed by two ranks")
            merged[k] = v
    return merged
