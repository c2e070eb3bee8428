"""Parity tests proper: the sm_100a kernels, called through the C-ABI, against
the oracle (bit-exact: reasons, fp64 priority bits, start/end times, node sets,
per-node core and gres slot masks) and against the committed golden fixtures;
plus size-independent invariants at sizes the oracle cannot finish."""
import os

import numpy as np
import pytest

from cranesched_b200 import abi, synth
from tests.helpers import assert_same, check_invariants, run_sched

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_config1_plumbing(oracle, gpu_lib):
    case = synth.config1()
    ref, _, _ = oracle.node_select(*case[:4], case[4])
    got, _ = run_sched(case, gpu_lib)
    assert_same(ref, got)


@pytest.mark.parametrize("seed", range(20, 32))
def test_random_feature_mix(oracle, gpu_lib, seed):
    """running jobs, fractional cpus, typed/untyped gres over two names, node
    lists, exclusive jobs, dead/drained nodes, unknown partitions, mandated
    priorities; multifactor and FIFO."""
    case = synth.random_case(seed, n_jobs=400, n_nodes=64, n_parts=1 + seed % 4, n_running=50,
                             fifo=bool(seed % 3 == 0))
    ref, _, _ = oracle.node_select(*case[:4], case[4])
    got, _ = run_sched(case, gpu_lib)
    assert_same(ref, got)
    check_invariants(case, got)


@pytest.mark.parametrize("seed", [40, 41, 42])
def test_timeline_cap_and_window(oracle, gpu_lib, seed):
    """kAlgoMaxJobNumPerNode cap (nodes drop out mid-tick) and the 7-day
    backfill window (JobScheduler.h:263-264, 809)."""
    case = synth.random_case(seed, n_jobs=600, n_nodes=12, n_parts=2, n_running=10, max_jobs_per_node=16)
    ref, _, _ = oracle.node_select(*case[:4], case[4])
    got, _ = run_sched(case, gpu_lib)
    assert_same(ref, got)


@pytest.mark.parametrize("seed", range(100, 124))
def test_random_sweep(oracle, gpu_lib, seed):
    """Small clusters of many shapes: partition sizes around the batch width (7
    nodes), one big partition, caps, short limits — the batch machinery of
    k_commit (selection lists, resolve, re-keying) under varied contention."""
    shapes = [dict(n_nodes=7, n_parts=1), dict(n_nodes=8, n_parts=1), dict(n_nodes=20, n_parts=2),
              dict(n_nodes=70, n_parts=1), dict(n_nodes=130, n_parts=3), dict(n_nodes=33, n_parts=4)]
    kw = dict(shapes[seed % len(shapes)])
    kw.update(n_jobs=250 + 37 * (seed % 11), n_running=seed % 40, fifo=bool(seed % 5 == 0), short=bool(seed % 2))
    if seed % 7 == 0:
        kw["max_jobs_per_node"] = 9
    case = synth.random_case(seed, **kw)
    ref, _, _ = oracle.node_select(*case[:4], case[4])
    got, _ = run_sched(case, gpu_lib)
    assert_same(ref, got)
    check_invariants(case, got)


@pytest.mark.parametrize("seed", range(200, 212))
def test_general_task_distribution(oracle, gpu_lib, seed):
    """ntasks_per_node_max > min and ntasks anywhere in [node_num*min, node_num*max]
    (JobScheduler.cpp:5193-5222, 5258-5361): per-node task counts, the top-K heaps
    with libstdc++'s tie behaviour, hand-out in pop order."""
    case = synth.random_case(seed, n_jobs=400, n_nodes=60, n_parts=1 + seed % 3, n_running=40, ntpn_range=True,
                             fifo=bool(seed % 4 == 0))
    ref, _, _ = oracle.node_select(*case[:4], case[4])
    got, _ = run_sched(case, gpu_lib)
    assert_same(ref, got)


@pytest.mark.parametrize("seed", range(300, 310))
def test_bestfit_policy(oracle, gpu_lib, seed):
    """cost_policy 1 (BestFit, BASELINE config 4): keys move DOWN when a node is
    allocated, so re-keyed nodes overtake their neighbours in the order."""
    case = synth.random_case(seed, n_jobs=400, n_nodes=60, n_parts=1 + seed % 3, n_running=40, cost_policy=1,
                             ntpn_range=bool(seed % 2), fifo=bool(seed % 4 == 0))
    ref, _, _ = oracle.node_select(*case[:4], case[4])
    got, _ = run_sched(case, gpu_lib)
    assert_same(ref, got)
    check_invariants(case, got)


@pytest.mark.parametrize("seed", range(500, 512))
def test_reservations(oracle, gpu_lib, seed):
    """Reservations (JobScheduler.cpp:5655-5753, 5788-5795, 5829-5861): schedulers of
    their own over the reserved resources, later reservations cut out of the
    timelines, running jobs inside reservations, the "Resource Reserved" and
    "Reservation Not Found" reasons."""
    case = synth.random_case(seed, n_jobs=400, n_nodes=60, n_parts=1 + seed % 3, n_running=30,
                             fifo=bool(seed % 4 == 0), ntpn_range=bool(seed % 3 == 0))
    resv, pd2, rn2 = synth.random_reservations(seed, case, n_resv=6)
    cfg, cl, rn, pd, now = case
    ref, _, _ = oracle.node_select(cfg, cl, rn2, pd2, now, resv=resv)
    got, _ = run_sched((cfg, cl, rn2, pd2, now), gpu_lib, resv=resv)
    assert_same(ref, got)


@pytest.mark.parametrize("seed", range(600, 612))
def test_overlapping_partitions(oracle, gpu_lib, seed):
    """Partitions that share nodes (one NodeState seen by several LocalSchedulers,
    JobScheduler.cpp:5597-5651): one scheduler per connected group, one order per
    partition; also with reservations on top and groups beside stand-alone partitions."""
    base = synth.random_case(seed, n_jobs=500, n_nodes=90, n_parts=2 + seed % 4, n_running=30,
                             fifo=bool(seed % 5 == 0), ntpn_range=bool(seed % 3 == 0), short=bool(seed & 1))
    case = synth.overlap_partitions(base, seed, frac=0.2 + 0.15 * (seed % 4), which=None if seed % 3 else {1})
    cfg, cl, rn, pd, now = case
    resv = None
    if seed % 4 == 1:
        resv, pd, rn = synth.random_reservations(seed, case, n_resv=4)
    ref, _, _ = oracle.node_select(cfg, cl, rn, pd, now, resv=resv)
    got, _ = run_sched((cfg, cl, rn, pd, now), gpu_lib, resv=resv)
    assert_same(ref, got)
    check_invariants((cfg, cl, rn, pd, now), got)


def test_batch_limit(oracle, gpu_lib):
    """ScheduledBatchSize: ranks beyond the limit get "Priority"."""
    case = synth.random_case(50, n_jobs=500, n_nodes=40, n_running=20, limit=137)
    ref, _, _ = oracle.node_select(*case[:4], case[4])
    got, _ = run_sched(case, gpu_lib)
    assert_same(ref, got)
    assert (got.reason == abi.REASON_PRIORITY).sum() >= 500 - 137


@pytest.mark.parametrize("name", ["random_7", "random_8", "config2_small"])
def test_golden_fixtures(gpu_lib, name):
    from tests.golden.make_golden import CASES
    case = CASES[name]()
    got, _ = run_sched(case, gpu_lib)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    ref = abi.Placements(**{f: g[f] for f in abi.Placements.__dataclass_fields__})
    assert_same(ref, got)


def test_config2_medium_vs_oracle(oracle, gpu_lib):
    """config-2 shape (4 partitions, cpu+mem+gres, multifactor + backfill) at a
    size the oracle finishes in seconds."""
    case = synth.config2(n_jobs=6000, n_nodes=600)
    ref, _, _ = oracle.node_select(*case[:4], case[4])
    got, _ = run_sched(case, gpu_lib)
    assert_same(ref, got)
    check_invariants(case, got)


def test_config5_backfill_stress_small(oracle, gpu_lib):
    case = synth.config5(n_jobs=3000, n_nodes=100)
    ref, _, _ = oracle.node_select(*case[:4], case[4])
    got, _ = run_sched(case, gpu_lib)
    assert_same(ref, got)


def test_config4_hetero_gres_small(oracle, gpu_lib):
    case = synth.config4(n_jobs=4000, n_nodes=250)
    ref, _, _ = oracle.node_select(*case[:4], case[4])
    got, _ = run_sched(case, gpu_lib)
    assert_same(ref, got)
    check_invariants(case, got)


def test_config2_full_size_invariants_and_prefix(oracle, gpu_lib):
    """BASELINE config 2 at full size (100k x 10k): invariants over the whole
    result, idempotence (same input -> same bits), and bit-exactness against
    the oracle on the prefix of the priority order the oracle can afford."""
    case = synth.config2()
    got, timing = run_sched(case, gpu_lib)
    check_invariants(case, got)
    again, _ = run_sched(case, gpu_lib)
    assert_same(got, again)
    cfg, cl, rn, pd, now = case
    n_pref = 3000
    ref, _, done = oracle.node_select(cfg, cl, rn, pd, now, max_jobs=n_pref)
    assert done == n_pref
    order = np.argsort(-ref.priority, kind="stable")[:n_pref]
    for f in ("reason", "start_time", "end_time", "n_alloc"):
        assert np.array_equal(getattr(ref, f)[order], getattr(got, f)[order]), f
    assert np.array_equal(ref.priority.view(np.uint64), got.priority.view(np.uint64))
    rep = np.repeat(np.arange(pd.n), pd.node_num)
    sel = np.isin(rep, order)
    assert np.array_equal(ref.alloc_node[sel], got.alloc_node[sel])
    assert ref.alloc_res[sel].tobytes() == got.alloc_res[sel].tobytes()


def test_capability_bitmap_rows(oracle, gpu_lib):
    """The jobs x nodes bitmap row of a job equals the oracle's per-(job,node)
    GetFeasibleResourceInNode(res_total) verdict."""
    from cranesched_b200.scheduler import GpuScheduler
    case = synth.random_case(60, n_jobs=120, n_nodes=40, n_parts=1, n_running=0, lists=False)
    cfg, cl, rn, pd, now = case
    s = GpuScheduler(cfg, 0, gpu_lib)
    s.set_cluster(cl)
    out = s.node_select(now, rn, pd)
    bm = s.debug_bitmap()
    s.close()
    usable = np.flatnonzero((cl.alive == 1) & (cl.drain == 0))
    queued = [j for j in np.argsort(-out.priority, kind="stable") if pd.partition[j] < cl.n_partitions]
    assert bm.shape[0] == len(queued)
    for r, j in enumerate(queued):
        req = np.zeros((), abi.RES_VIEW)
        t = int(pd.ntasks_per_node_min[j])
        for f in ("cpu_raw", "mem", "mem_sw"):
            req[f] = int(pd.req_node[j][f]) + int(pd.req_task[j][f]) * t
        req["gres_total"] = pd.req_node[j]["gres_total"] + pd.req_task[j]["gres_total"] * t
        req["gres_spec"] = pd.req_node[j]["gres_spec"] + pd.req_task[j]["gres_spec"] * t
        for q, node in enumerate(usable):
            want = oracle.feasible(cl, req, cl.res_total[node])[0]
            have = bool(bm[r, q // 32] >> (q % 32) & 1)
            assert want == have, (r, j, q)
