import numpy as np

from cranesched_b200 import abi
from cranesched_b200.scheduler import GpuScheduler


def run_sched(case, lib_path=None, device=0, resv=None):
    cfg, cluster, running, pending, now = case
    s = GpuScheduler(cfg, device, lib_path)
    try:
        s.set_cluster(cluster)
        if resv is not None:
            s.set_reservations(resv)
        out = s.node_select(now, running, pending)
        timing = s.timing()
    finally:
        s.close()
    return out, timing


def assert_same(ref: abi.Placements, got: abi.Placements):
    d = ref.diff(got)
    assert not d, "placements differ from the oracle:\n" + "\n".join(d[:12])


def check_invariants(case, out: abi.Placements):
    """Size-independent properties of a NodeSelect result (used at full scale,
    where the oracle cannot finish): no node is over-committed at any time and
    every placement is internally consistent."""
    cfg, cluster, running, pending, now = case
    placed = out.n_alloc > 0
    assert ((out.reason == abi.REASON_NONE) <= placed).all()
    assert (out.n_alloc[placed] == pending.node_num[placed]).all()
    assert (out.start_time[placed] >= now).all()
    assert (out.end_time[placed] - out.start_time[placed] == pending.time_limit[placed]).all()
    assert (out.start_time[out.reason == abi.REASON_NONE] == now).all()
    started = np.flatnonzero(out.reason == abi.REASON_NONE)
    # resources of jobs started now never exceed node totals
    used_cpu = np.zeros(cluster.n_nodes, np.int64)
    used_mem = np.zeros(cluster.n_nodes, np.float64)
    cores = np.zeros((cluster.n_nodes, abi.CORE_WORDS), np.uint64)
    gres = np.zeros((cluster.n_nodes, abi.GRES_ENTRIES), np.uint16)
    rep = np.repeat(np.arange(pending.n), pending.node_num)
    sel = np.isin(rep, started)
    nodes = out.alloc_node[sel]
    res = out.alloc_res[sel]
    np.add.at(used_cpu, nodes, res["cpu_raw"])
    np.add.at(used_mem, nodes, res["mem"].astype(np.float64))
    assert (used_cpu <= cluster.res_total["cpu_raw"]).all()
    assert (used_mem <= cluster.res_total["mem"].astype(np.float64) + 1).all()
    # concrete cores / slots handed out at `now` are pairwise disjoint per node
    order = np.argsort(nodes, kind="stable")
    for idx in order:
        n = nodes[idx]
        assert not (cores[n] & res["core"][idx]).any(), f"core double-booked on node {n}"
        assert not (gres[n] & res["gres"][idx]).any(), f"gres slot double-booked on node {n}"
        cores[n] |= res["core"][idx]
        gres[n] |= res["gres"][idx]
        assert not (res["core"][idx] & ~cluster.res_total["core"][n]).any()
        assert not (res["gres"][idx] & ~cluster.res_total["gres"][n]).any()
