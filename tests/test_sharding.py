"""Multi-GPU plumbing on CPU: the bench shards instances by rank with no data-path collective and
reduces only timings (max) and FLOP counters (sum).  world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scenario import Scenario


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B = 8
    sc = Scenario(B, gaits="trot", seed=20260 + rank)            # bench.py: one seed stream per rank
    xref, fsteps = sc.inputs()
    t = torch.tensor([0.010 * (rank + 1)], dtype=torch.float64)  # pretend elapsed seconds
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([float(B)], dtype=torch.float64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    gathered = [torch.zeros(B, 12, 17, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(xref))
    if rank == 0:
        out.put((float(t.item()), float(n.item()), bool(torch.equal(gathered[0], gathered[1]))))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    tmax, total, same = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert abs(tmax - 0.020) < 1e-12          # slowest rank defines the step time
    assert total == 16.0                      # whole-job instances = sum over ranks
    assert not same                           # ranks work on different robots (disjoint shards)


def test_shards_are_independent_of_batch_position():
    """An instance's inputs depend only on its own seed/gait/command, not on who shares the batch."""
    big = Scenario(6, gaits=["trot", "pace"], seed=3)
    xr, fs = big.inputs()
    sub = Scenario(1, gaits=big.kinds[4], v_ref=big.v_ref[4], phase=[big.phase[4]], random_commands=False)
    x1, f1 = sub.inputs()
    np.testing.assert_array_equal(x1[0], xr[4])
