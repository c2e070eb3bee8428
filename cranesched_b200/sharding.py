"""One scheduling tick over several GPUs.

The job loop of NodeSelect is independent per partition — one LocalScheduler
each, with its own node set (JobScheduler.cpp:5757-5766; "TODO: do it in
parallel", :5756) — so ONE queue is split by dealing its partitions to the
ranks (largest job count first). Every rank uploads the whole pending table
(priority, batch limit and queue order are global: JobScheduler.cpp:6545-6671),
commits only its own partitions, zeroes the placement columns of the jobs it
does not own, and the union over ranks is one all-reduce(sum) per column —
NCCL over NVLink on GPUs, gloo in the CPU tests. The result on every rank is
bit-identical to the single-GPU tick.

Inside one partition the loop is a dependency chain of ~20 microsecond batches
(DESIGN.md §5): there is nothing left to hand to another GPU at NVLink latency,
so partitions are the unit and a queue with fewer partitions than ranks leaves
ranks idle. `shard_workload` is the other line of the bench: independent
clusters, one per rank (weak scaling).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi, synth


def shard_workload(config_id: int, rank: int, world: int, n_jobs: int = 0, n_nodes: int = 0):
    """Weak scaling: rank r schedules its own cluster (config-shaped, its own
    draw: seed = 1000 * rank + config) — `world` independent clusters."""
    gen = synth.CONFIGS[config_id]
    kw = {}
    if n_jobs:
        kw["n_jobs"] = n_jobs
    if n_nodes:
        kw["n_nodes"] = n_nodes
    if config_id != 1:
        kw["seed_id"] = config_id + 1000 * rank
    return gen(**kw)


def partition_groups(cluster: abi.Cluster) -> np.ndarray:
    """partition -> group id; partitions that share a node (directly or through a
    chain of partitions) are one group: they are one scheduler of the library and
    must have one owner."""
    n = cluster.n_partitions
    parent = list(range(n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    first = {}
    for p in range(n):
        for node in cluster.part_nodes[cluster.part_off[p]:cluster.part_off[p + 1]].tolist():
            q = first.setdefault(node, p)
            a, b = find(q), find(p)
            if a != b:
                parent[max(a, b)] = min(a, b)
    return np.array([find(p) for p in range(n)], np.uint32)


def deal_partitions(pending: abi.Pending, n_partitions: int, world: int, cluster: abi.Cluster = None) -> np.ndarray:
    """partition -> rank, longest processing time first on the job counts (the
    tick of a rank is the longest job loop among its partitions, and one GPU
    runs its partitions concurrently, so what matters is to spread the big ones).
    With `cluster`, groups of overlapping partitions are dealt as one."""
    njobs = np.bincount(pending.partition[pending.partition < n_partitions], minlength=n_partitions)
    group = partition_groups(cluster) if cluster is not None else np.arange(n_partitions, dtype=np.uint32)
    gjobs = np.bincount(group, weights=njobs, minlength=n_partitions)
    load = np.zeros(world, np.int64)
    gowner = np.zeros(n_partitions, np.uint32)
    for g in np.argsort(-gjobs, kind="stable"):
        if not (group == g).any():
            continue
        r = int(np.argmin(load))
        gowner[g] = r
        load[r] += int(gjobs[g])
    return gowner[group].astype(np.uint32)


class _DevArray:
    """Zero-copy view of a device buffer of the C-ABI library for torch
    (`__cuda_array_interface__`)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _columns(sched, on_gpu: bool):
    """The placement columns of the last run as torch tensors aliasing the
    library's buffers (device memory on a GPU, host memory under the CPU
    kernel-emulation harness of the tests)."""
    import torch
    d = sched.device_placements()
    nj, nr = int(d.n_jobs), int(d.n_rows)
    spec = [(d.reason, nj, "|u1", np.uint8), (d.start_time, nj, "<i8", np.int64), (d.end_time, nj, "<i8", np.int64),
            (d.n_alloc, nj, "<i4", np.int32), (d.alloc_node, nr, "<i4", np.int32), (d.alloc_ntasks, nr, "<i4", np.int32),
            (d.alloc_res, nr * 9, "<i8", np.int64)]
    out = []
    for ptr, n, ts, dt in spec:
        if n == 0 or not ptr:
            continue
        if on_gpu:
            out.append(torch.as_tensor(_DevArray(ptr, n, ts), device="cuda"))
        else:
            buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
            out.append(torch.from_numpy(np.frombuffer(buf, dtype=dt)))
    return out


def sharded_tick(sched, now: int, running: abi.Running, pending: abi.Pending, owner: np.ndarray, dist,
                 out: abi.Placements | None = None, on_gpu: bool = True, upload: bool = True) -> abi.Placements:
    """One NodeSelect over dist.get_world_size() GPUs: upload (whole tables), run
    (own partitions), all-reduce of the placement columns, fetch. Every rank
    returns the whole tick."""
    rank, world = dist.get_rank(), dist.get_world_size()
    sched.set_shard(rank, world, owner)
    if upload:
        sched.upload(running, pending)
    sched.run(now)
    sched.sync()  # the library's stream is not torch's: the columns are complete before the collective reads them
    cols = _columns(sched, on_gpu)
    for t in cols:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if on_gpu:
        import torch
        torch.cuda.synchronize()
    out = out if out is not None else abi.Placements.for_pending(pending)
    return sched.fetch(out)


def reduce_metric(decided_local: int, device_ms_local: float, dist=None, device="cpu"):
    """Whole-job aggregate: decisions of all ranks / max device time over ranks."""
    import torch
    t = torch.tensor([float(device_ms_local)], dtype=torch.float64, device=device)
    n = torch.tensor([float(decided_local)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(n.item()), float(t.item())
