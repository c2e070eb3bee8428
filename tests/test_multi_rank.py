"""N>1 path on CPU: world_size-2 gloo processes run ONE queue split by partition
(crane_sched_set_shard + all-reduce of the placement columns) through the
kernel-emulation library; every rank must end up with the whole tick, bit-equal
to the oracle's unsplit answer — over several seeds, with running jobs and with
a batch limit smaller than the queue (priority, fair-share and the cut-off are
global: JobScheduler.cpp:6545-6671). Also the weak-scaling shard (independent
clusters per rank) and the aggregate-metric reduction bench.py uses."""
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
    from cranesched_b200.scheduler import GpuScheduler
    from oracle import pyoracle
    from tests.helpers import run_sched

    # (1) weak scaling: every rank schedules its own cluster (its own draw)
    case = sharding.shard_workload(2, rank, world, n_jobs=90, n_nodes=16)
    got, _ = run_sched(case, emu_lib)
    ref, _, _ = pyoracle.node_select(*case[:4], case[4])
    ok_weak = not ref.diff(got)
    total, tmax = sharding.reduce_metric(case[3].n, 10.0 * (rank + 1), dist)
    seed_sig = torch.tensor([int(case[3].time_limit[:8].sum())])
    gathered = [torch.zeros_like(seed_sig) for _ in range(world)]
    dist.all_gather(gathered, seed_sig)

    # (2) ONE queue over the ranks: partitions dealt out, columns all-reduced
    bad = []
    cases = [synth.config2(n_jobs=90, n_nodes=20, seed_id=77),
             synth.random_case(32, n_jobs=90, n_nodes=30, n_parts=4, n_running=16, limit=60),
             synth.random_case(33, n_jobs=70, n_nodes=18, n_parts=1, n_running=5),     # fewer partitions than ranks
             synth.random_case(35, n_jobs=80, n_nodes=27, n_parts=3, n_running=12, fifo=True),
             # partitions 1 and 2 share nodes: one scheduler, one owner; 0 and 3 stand alone
             synth.overlap_partitions(synth.random_case(36, n_jobs=90, n_nodes=32, n_parts=4, n_running=10), 36, 0.5, which={2})]
    for k, full in enumerate(cases):
        cfg, cl, rn, pd, now = full
        owner = sharding.deal_partitions(pd, cl.n_partitions, world, cl)
        s = GpuScheduler(cfg, 0, emu_lib)
        s.set_cluster(cl)
        got = sharding.sharded_tick(s, now, rn, pd, owner, dist, on_gpu=False)
        s.close()
        ref_full, _, _ = pyoracle.node_select(cfg, cl, rn, pd, now)
        d = ref_full.diff(got)
        if d:
            bad.append((k, d[:3]))
    q.put((rank, ok_weak, bad, total, tmax, [int(g.item()) for g in gathered]))
    dist.destroy_process_group()


def test_two_ranks_gloo(oracle, emu_lib):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, emu_lib, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_weak, bad, total, tmax, seeds in res:
        assert ok_weak, f"rank {rank}: its own cluster's result differs from the oracle"
        assert not bad, f"rank {rank}: the sharded tick differs from the unsplit answer: {bad}"
        assert total == 180.0 and tmax == 20.0  # sum of decisions, max of times
        assert seeds[0] != seeds[1]  # weak scaling: every rank has its own draw
