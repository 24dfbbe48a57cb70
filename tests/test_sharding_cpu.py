"""N > 1 host logic on CPU (gloo, world_size 2): independent sequences shard over ranks with no data-path collective;
every sequence is processed exactly once and gives the single-process result bit for bit (the oracle stands in for
the engine here -- no GPU in this container)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.common import ROOT, make_oracle, synth
from msckf_mono_b200 import shard

N_SEQ = 5


def run_sequence(seq):
    wl = synth.make_window_workload(n_features=10, n_clones=6, seq=seq)
    o = make_oracle(ROOT / "oracle" / "libmsckf_oracle.so", np.float64)
    synth.drive(o, wl)
    return {"p": o.getImuState()["p_I_G"].tolist(), "trace": float(np.trace(o.getCovariance())), "acc": int(o.lastReport()["accepted"].sum())}


def worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.sequences_of_rank(N_SEQ, rank, world)
    local = {s: run_sequence(s) for s in mine}
    shard.barrier()
    merged = shard.gather_summaries(local)
    tmax = shard.max_over_ranks([float(rank + 1), 10.0 - rank])
    if rank == 0:
        q.put((merged, tmax, mine))
    dist.destroy_process_group()


def test_two_ranks_cover_all_sequences_once(oracle_lib):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, tmax, mine0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(merged) == list(range(N_SEQ)) and mine0 == [0, 2, 4]
    assert tmax == [2.0, 10.0]
    for s in range(N_SEQ):
        assert merged[s] == run_sequence(s)  # identical to the single-process result


def test_round_robin_partition():
    for world in (1, 2, 4, 8):
        allseq = sorted(s for r in range(world) for s in shard.sequences_of_rank(64, r, world))
        assert allseq == list(range(64))
        assert max(len(shard.sequences_of_rank(64, r, world)) for r in range(world)) == 64 // world
