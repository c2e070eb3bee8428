"""QoS post-filter: CheckAndMallocQosResource over the
placements NodeSelect produced (JobScheduler.cpp:1262;
Accounting/AccountMetaContainer.cpp:164-191, 382-531, 546-587).

CPU part: hand-derived known answers for the oracle restatement + the real
kernel source under the CUDA emulation. GPU part: parity through the C-ABI."""
import numpy as np
import pytest

from cranesched_b200 import abi, synth
from cranesched_b200.scheduler import GpuScheduler


def _tiny():
    """4 single-node jobs, 1 partition of 4 big nodes, everything placed now."""
    cfg, cluster, running, pend, now = synth.config1(n_jobs=4, n_nodes=4)
    pend.qos = np.zeros(4, np.uint32)  # fresh arrays: config1 shares one zeros column
    pend.user = np.array([0, 0, 1, 1], np.uint32)
    pend.account = np.array([0, 0, 1, 1], np.uint32)
    for col in ("req_node", "req_task", "req_total", "node_num", "ntasks", "ntasks_per_node_min",
                "ntasks_per_node_max", "exclusive"):
        getattr(pend, col)[:] = getattr(pend, col)[0]  # every job asks the same
    return cfg, cluster, running, pend, now


def _table(pend, **kw):
    big = 1 << 60
    q = 1
    unl = np.zeros(q, abi.TRES_LIMIT)
    unl["view"]["cpu_raw"] = big
    unl["view"]["mem"] = big
    d = dict(
        n_users=2, n_accounts=3, valid=[1], max_jobs_per_user=[1000], max_jobs_per_account=[1000],
        max_jobs=[1000], max_cpus_per_user_raw=[big], max_wall=[0], max_tres_per_user=unl.copy(),
        max_tres_per_account=unl.copy(), max_tres=unl.copy(),
        chain_off=np.arange(pend.n + 1) * 2,
        chain_acct=np.stack([pend.account, np.full(pend.n, 2)], 1).reshape(-1),
    )
    d.update(kw)
    return abi.QosTable(**d)


def _oracle_filter(oracle, case, table):
    out, _, _ = oracle.node_select(*case[:4], case[4])
    assert (out.reason == 0).all(), "fixture expects every job to start now"
    oracle.qos_filter(case[1], case[3], out, table)
    return out


def test_oracle_known_answers(oracle):
    case = _tiny()
    pend = case[3]
    cpu = int(pend.req_total["cpu_raw"][0])  # every job asks the same
    # 1. unlimited: all pass, usage = sum of allocations at every level
    t = _table(pend)
    out = _oracle_filter(oracle, case, t)
    assert out.reason.tolist() == [0, 0, 0, 0]
    assert t.user_usage["jobs_count"].tolist() == [2, 2]
    assert t.account_usage["jobs_count"].tolist() == [2, 2, 4]
    assert t.qos_usage["jobs_count"].tolist() == [4]
    assert t.qos_usage["cpu_raw"][0] == 4 * cpu
    assert t.qos_usage["wall_time"][0] == int(pend.time_limit.sum())
    # 2. max_jobs_per_user = 1: second job of each user fails with Jobs
    t = _table(pend, max_jobs_per_user=[1])
    out = _oracle_filter(oracle, case, t)
    assert out.reason.tolist() == [0, abi.REASON_QOS_JOBS, 0, abi.REASON_QOS_JOBS]
    assert t.qos_usage["jobs_count"].tolist() == [2]
    # 3. max_cpus_per_user: the cpu check precedes the jobs check
    t = _table(pend, max_jobs_per_user=[1], max_cpus_per_user_raw=[2 * cpu - 1])
    out = _oracle_filter(oracle, case, t)
    assert out.reason.tolist() == [0, abi.REASON_QOS_CPU, 0, abi.REASON_QOS_CPU]
    # 4. the root account limits the whole chain: 3 jobs in total
    t = _table(pend, max_jobs_per_account=[3])
    out = _oracle_filter(oracle, case, t)
    assert out.reason.tolist() == [0, 0, 0, abi.REASON_QOS_JOBS]
    assert t.account_usage["jobs_count"].tolist() == [2, 1, 3]
    # 5. wall time: limit reached by the qos total
    tl = pend.time_limit
    t = _table(pend, max_wall=[int(tl[0] + tl[1] + tl[2])])
    out = _oracle_filter(oracle, case, t)
    assert out.reason.tolist() == [0, 0, 0, abi.REASON_QOS_WALL]
    # 6. max_tres memory at the qos level
    lim = _table(pend).max_tres.copy()
    lim["view"]["mem"] = int(pend.req_total["mem"][0]) * 2
    t = _table(pend, max_tres=lim)
    out = _oracle_filter(oracle, case, t)
    assert out.reason.tolist() == [0, 0, abi.REASON_QOS_MEM, abi.REASON_QOS_MEM]
    # 7. deleted qos
    t = _table(pend, valid=[0])
    out = _oracle_filter(oracle, case, t)
    assert out.reason.tolist() == [abi.REASON_QOS_INVALID] * 4
    assert t.qos_usage["jobs_count"].tolist() == [0]


def test_oracle_gres_limit_semantics(oracle):
    """CheckGres_ (AccountMetaContainer.cpp:509-531): a name / type missing from
    the limit map ends the check with "pass"."""
    cfg, cluster, running, pend, now = synth.config3(n_jobs=40, n_nodes=64, n_parts=4)
    out, _, _ = oracle.node_select(cfg, cluster, running, pend, now)
    started = out.reason == 0
    gpu_jobs = np.flatnonzero(started & (pend.req_total["gres_total"][:, 0] > 0))
    assert len(gpu_jobs) >= 2
    pend.qos = np.zeros(pend.n, np.uint32)
    base = synth.random_qos(1, cluster, pend, tight=1e6, invalid_frac=0.0)
    base.user_usage[:] = 0
    base.account_usage[:] = 0
    base.qos_usage[:] = 0

    def run(name_present, spec_present, total, spec):
        t = base.copy()
        t.max_tres["gres_name_present"] = name_present
        t.max_tres["gres_spec_present"] = spec_present
        t.max_tres["view"]["gres_total"][:] = total
        t.max_tres["view"]["gres_spec"][:] = spec
        o = abi.Placements(**{f: getattr(out, f).copy() for f in out.__dataclass_fields__})
        oracle.qos_filter(cluster, pend, o, t)
        return o.reason

    # name absent from the limit: unlimited even with zero counts
    assert (run(0, 0, 0, 0)[gpu_jobs] == 0).all()
    # name present, total 0: every gpu job fails, the others pass
    r = run(1, 0, 0, 0)
    assert (r[gpu_jobs] == abi.REASON_QOS_GRES).all()
    assert (r[started & (pend.req_total["gres_total"][:, 0] == 0)] == 0).all()
    # total generous, types present with 0: every gpu job fails on its type
    r = run(1, 0xFF, 60000, 0)
    assert (r[gpu_jobs] == abi.REASON_QOS_GRES).all()
    # total generous, types absent: pass
    assert (run(1, 0, 60000, 0)[gpu_jobs] == 0).all()


def _parity(oracle, lib_path, case, table, device=0):
    cfg, cluster, running, pend, now = case
    ref, _, _ = oracle.node_select(cfg, cluster, running, pend, now)
    t_ref = table.copy()
    oracle.qos_filter(cluster, pend, ref, t_ref)
    t_got = table.copy()
    s = GpuScheduler(cfg, device, lib_path)
    try:
        s.set_cluster(cluster)
        got = s.node_select(now, running, pend)
        s.qos_filter(t_got, got.reason)
    finally:
        s.close()
    assert not ref.diff(got), ref.diff(got)[:5]
    for f in ("user_usage", "account_usage", "qos_usage"):
        a, b = getattr(t_ref, f), getattr(t_got, f)
        assert a.tobytes() == b.tobytes(), f
    return ref


@pytest.mark.parametrize("in_global", [False, True])
def test_emulated_kernel_matches_oracle(oracle, emu_lib, monkeypatch, in_global):
    if in_global:
        monkeypatch.setenv("CRANE_QOS_TABLES_GLOBAL", "1")
    # more than one prepared batch (96 jobs) per qos
    case = synth.random_case(21, n_jobs=150 if in_global else 300, n_nodes=60, n_running=8)
    table = synth.random_qos(21, case[1], case[3], tight=1.0)
    ref = _parity(oracle, emu_lib, case, table)
    assert (ref.reason >= 16).any()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,tight,invalid", [(31, 1.0, 0.0), (32, 3.0, 0.1), (33, 0.5, 0.4)])
def test_gpu_matches_oracle(oracle, seed, tight, invalid):
    case = synth.random_case(seed, n_jobs=1500, n_nodes=160, n_parts=4, n_running=60)
    table = synth.random_qos(seed, case[1], case[3], tight=tight, invalid_frac=invalid)
    ref = _parity(oracle, None, case, table)
    codes = set(np.unique(ref.reason).tolist())
    assert codes & {16, 17, 18, 19, 20, 21}


@pytest.mark.gpu
def test_gpu_usage_tables_in_global_memory(oracle, monkeypatch):
    """The path of usage tables too large for shared memory (forced here)."""
    monkeypatch.setenv("CRANE_QOS_TABLES_GLOBAL", "1")
    case = synth.random_case(34, n_jobs=1500, n_nodes=160, n_parts=4, n_running=60)
    table = synth.random_qos(34, case[1], case[3], tight=1.5, invalid_frac=0.1)
    ref = _parity(oracle, None, case, table)
    assert set(np.unique(ref.reason).tolist()) & {16, 17, 18, 19, 20, 21}


@pytest.mark.gpu
def test_gpu_config3_slice(oracle):
    """Config 3 (the configuration BASELINE.json names for the QoS filter) at a
    size the oracle finishes in seconds."""
    case = synth.config3(n_jobs=6000, n_nodes=800, n_parts=8)
    table = synth.random_qos(3, case[1], case[3], tight=8.0)
    ref = _parity(oracle, None, case, table)
    assert (ref.reason == 0).any() and (ref.reason >= 16).any()
