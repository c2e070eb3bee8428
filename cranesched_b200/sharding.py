"""Multi-GPU layout of the scheduling tick.

The path shards by partition: every partition has its own LocalScheduler and
its own node set (JobScheduler.cpp:5757-5766), so disjoint partitions never
interact inside a tick. Rank r of N therefore owns a disjoint slice of the
cluster's partitions together with the pending jobs of those partitions and
runs the unmodified single-GPU tick on it; there is no data-path collective.
`torch.distributed` (nccl on GPUs, gloo in the CPU tests) carries the barrier,
the max-over-ranks timing and the gather of per-rank summaries only.
"""
from __future__ import annotations

import numpy as np

from . import abi, synth


def shard_workload(config_id: int, rank: int, world: int, n_jobs: int = 0, n_nodes: int = 0):
    """Weak-scaling shard: rank r gets its own config-shaped set of partitions (a
    disjoint slice of a `world`-times larger cluster). Every rank's slice is
    generated from the same seed, so the per-GPU work is the same for every N —
    the definition of weak scaling; with per-rank seeds the max over ranks
    measures the spread of the synthetic workloads instead (one of eight
    config-2 draws takes 1.8x the cycles of the others)."""
    gen = synth.CONFIGS[config_id]
    kw = {}
    if n_jobs:
        kw["n_jobs"] = n_jobs
    if n_nodes:
        kw["n_nodes"] = n_nodes
    if config_id != 1:
        kw["seed_id"] = config_id
    return gen(**kw)


def split_by_partition(case, world: int):
    """Strong split of ONE cluster: partitions are dealt to ranks (largest job
    count first, LPT), each rank keeps the nodes and jobs of its partitions.
    Returns per-rank (case, job_index) so results can be scattered back."""
    cfg, cl, rn, pd, now = case
    njobs = np.bincount(pd.partition[pd.partition < cl.n_partitions], minlength=cl.n_partitions)
    load = np.zeros(world, np.int64)
    owner = np.zeros(cl.n_partitions, np.int64)
    for p in np.argsort(-njobs, kind="stable"):
        r = int(np.argmin(load))
        owner[p] = r
        load[r] += njobs[p]
    out = []
    for r in range(world):
        parts = np.flatnonzero(owner == r)
        node_sel = np.concatenate([cl.part_nodes[cl.part_off[p]:cl.part_off[p + 1]] for p in parts]) if len(parts) else np.zeros(0, np.uint32)
        remap = np.full(cl.n_nodes, -1, np.int64)
        remap[node_sel] = np.arange(len(node_sel))
        off = [0]
        for p in parts:
            off.append(off[-1] + int(cl.part_off[p + 1] - cl.part_off[p]))
        sub_cl = abi.Cluster(cl.res_total[node_sel], cl.alive[node_sel], cl.drain[node_sel],
                             np.array(off, np.uint32), np.arange(len(node_sel), dtype=np.uint32),
                             cl.n_gres_entries, cl.gres_entry_name)
        pmap = np.full(cl.n_partitions + 1, len(parts), np.int64)  # unknown partitions stay unknown
        pmap[parts] = np.arange(len(parts))
        jsel = np.flatnonzero(np.isin(pd.partition, parts) | ((pd.partition >= cl.n_partitions) & (r == 0)))
        cols = {}
        for f in pd.__dataclass_fields__:
            v = getattr(pd, f)
            if f in ("incl_off", "incl_nodes", "excl_off", "excl_nodes") or v is None:
                continue
            cols[f] = v[jsel]
        cols["partition"] = pmap[np.minimum(cols["partition"], cl.n_partitions)].astype(np.uint32)
        for k in ("incl", "excl"):
            o = getattr(pd, k + "_off")
            if o is not None:
                nodes = getattr(pd, k + "_nodes")
                new_off, new_nodes = [0], []
                for j in jsel:
                    lst = remap[nodes[o[j]:o[j + 1]]]
                    new_nodes += lst[lst >= 0].tolist()
                    new_off.append(len(new_nodes))
                cols[k + "_off"] = np.array(new_off, np.uint32)
                cols[k + "_nodes"] = np.array(new_nodes, np.uint32)
        sub_pd = abi.Pending(**cols)
        if rn.n:
            raise NotImplementedError("split_by_partition: running jobs are not split yet")
        out.append(((cfg, sub_cl, rn, sub_pd, now), jsel, node_sel))
    return out


def reduce_metric(decided_local: int, device_ms_local: float, dist=None, device="cpu"):
    """Whole-job aggregate: decisions of all ranks / max device time over ranks."""
    import torch
    t = torch.tensor([float(device_ms_local)], dtype=torch.float64, device=device)
    n = torch.tensor([float(decided_local)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(n.item()), float(t.item())
