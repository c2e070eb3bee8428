"""N>1 path on CPU: world_size-2 gloo processes, each running the tick on its
shard through the kernel-emulation library, checked against the oracle and
against the aggregate-metric reduction bench.py uses."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, emu_lib, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cranesched_b200 import sharding, synth
    from oracle import pyoracle
    from tests.helpers import run_sched

    # (1) weak scaling: every rank owns its own set of partitions
    case = sharding.shard_workload(2, rank, world, n_jobs=90, n_nodes=16)
    got, _ = run_sched(case, emu_lib)
    ref, _, _ = pyoracle.node_select(*case[:4], case[4])
    ok_weak = not ref.diff(got)
    total, tmax = sharding.reduce_metric(case[3].n, 10.0 * (rank + 1), dist)
    # (2) one cluster split by partition: results scatter back to the unsplit answer
    full = synth.config2(n_jobs=120, n_nodes=24, seed_id=77)
    mine, jsel, node_sel = sharding.split_by_partition(full, world)[rank]
    sub, _ = run_sched(mine, emu_lib)
    ref_full, _, _ = pyoracle.node_select(*full[:4], full[4])
    ok_split = bool(np.array_equal(sub.reason, ref_full.reason[jsel])
                    and np.array_equal(sub.start_time, ref_full.start_time[jsel])
                    and np.array_equal(sub.n_alloc, ref_full.n_alloc[jsel]))
    # allocated nodes map back through node_sel
    rep_sub = np.repeat(np.arange(mine[3].n), mine[3].node_num)
    placed = sub.n_alloc[rep_sub] > 0
    rep_full = np.repeat(np.arange(full[3].n), full[3].node_num)
    full_rows = np.flatnonzero(np.isin(rep_full, jsel))
    ok_split = ok_split and bool(np.array_equal(node_sel[sub.alloc_node[placed]], ref_full.alloc_node[full_rows][placed]))
    seeds_differ = torch.tensor([int(case[3].time_limit[:8].sum())])
    gathered = [torch.zeros_like(seeds_differ) for _ in range(world)]
    dist.all_gather(gathered, seeds_differ)
    q.put((rank, ok_weak, ok_split, total, tmax, [int(g.item()) for g in gathered]))
    dist.destroy_process_group()


def test_two_ranks_gloo(oracle, emu_lib):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, emu_lib, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_weak, ok_split, total, tmax, seeds in res:
        assert ok_weak, f"rank {rank}: shard result differs from the oracle"
        assert ok_split, f"rank {rank}: partition split does not reproduce the unsplit answer"
        assert total == 180.0 and tmax == 20.0  # sum of decisions, max of times
        assert seeds[0] == seeds[1]  # weak scaling: every rank gets the same draw (equal per-GPU work)
