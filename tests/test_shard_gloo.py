"""CPU, world_size 2 on gloo: the subgroup sharding used for multi-GPU runs (rambl_b200/shard.py).
The per-rank solver is a stand-in here (the oracle is the checker on the CPU box, never the product);
what is tested is the host logic: every subgroup is solved exactly once, by one rank, and rank 0 gets
the results back in the original order."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from rambl_b200 import shard, synth  # noqa: E402


def test_assignment_is_a_partition_and_balanced():
    costs = [5.0, 1.0, 9.0, 3.0, 3.0, 7.0, 2.0]
    for world in (1, 2, 3, 8):
        parts = shard.assign(costs, world)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(len(costs)))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) <= sum(costs) / world + max(costs)
    assert shard.assign(costs, 2) == shard.assign(costs, 2)
    assert shard.assign([], 4) == [[], [], [], []]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sgs = [synth.make_subgroup(40 + 10 * k, 30, 2, seed=k, window=(50 * k, 50 * k + 80)) for k in range(7)]
    seen = []

    def solve(mine):
        seen.extend(s.gene for s in mine)
        return [">%s\n%d\n" % (s.gene[:12], s.n_reads) for s in mine]

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    res = shard.solve_sharded(sgs, rank, world, solve, gather)
    counts = [None] * world
    dist.all_gather_object(counts, len(seen))
    if rank == 0:
        q.put((res, counts, [">%s\n%d\n" % (s.gene[:12], s.n_reads) for s in sgs]))
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_solve_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res, counts, want = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == want
    assert sum(counts) == 7 and all(c > 0 for c in counts)
