"""world_size-2 `gloo` test of the multi-GPU path's host logic (runs on CPU): utterance shards
are disjoint and cover the sweep, per-rank step batches differ, and the job-level reduction
(max time, summed work) that bench.py reports is what every rank sees."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench
        from avsl_b200 import frontend, synth
        idx, durs = bench.rank_utterances(rank, world, 32)
        full = frontend.shard(bench.N_SWEEP, rank, world)
        gathered = [None] * world
        dist.all_gather_object(gathered, full.tolist())
        step_idx = [None] * world
        dist.all_gather_object(step_idx, idx.tolist())
        # the length-balanced shards: every rank derives the same partition on its own
        alld = synth.ami_durations(bench.N_SWEEP, bench.SEED)
        bal = [None] * world
        dist.all_gather_object(bal, frontend.shard_balanced(alld, rank, world).tolist())
        bal_loads = [float(alld[np.asarray(x, dtype=np.int64)].sum()) for x in bal]
        # per-rank fake timing: rank r took (10 + r) ms for its own audio seconds
        t, a, b, l = frontend.aggregate_rank_stats(10.0 + rank, float(durs.sum()), 1000.0 * (rank + 1), 5.0)
        if rank == 0:
            out.put({"cover": sorted(sum(gathered, [])) == list(range(bench.N_SWEEP)),
                     "disjoint": len(set(step_idx[0]) & set(step_idx[1])) == 0,
                     "t": t, "audio": a, "bytes": b, "launches": l,
                     "bal_cover": sorted(sum(bal, [])) == list(range(bench.N_SWEEP)),
                     "bal_gap": abs(bal_loads[0] - bal_loads[1]), "max_dur": float(alld.max()),
                     "expect_audio": float(synth.ami_durations(bench.N_SWEEP, bench.SEED)[np.r_[0:64]].sum())})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_sharding_and_reduction():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=100)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert res["cover"] and res["disjoint"]
    assert res["bal_cover"] and res["bal_gap"] <= res["max_dur"]
    assert res["t"] == 11.0                       # max over ranks
    assert res["bytes"] == 3000.0 and res["launches"] == 10.0
    # ranks 0 and 1 own utterances 0,2,..,62 and 1,3,..,63: together the first 64 of the sweep
    assert abs(res["audio"] - res["expect_audio"]) < 1e-9
