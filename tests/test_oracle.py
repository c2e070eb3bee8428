"""Oracle checks (CPU): the reference's own known-answer vectors, golden
fixtures, and the bit-mask algebra the kernels use vs the container algebra."""
import os

import numpy as np
import pytest

from cranesched_b200 import abi, synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_reference_kats(oracle):
    """test/Utilities/dedicated_resource_test.cpp:27-171 (14 cases) + micro-cases."""
    failed, log = oracle.selftest()
    assert failed == 0, log


def test_config1_all_start_now(oracle):
    """Config 1 (1k x 128, FIFO): the queue fits, every job starts now."""
    cfg, cl, rn, pd, now = synth.config1()
    out, ms, done = oracle.node_select(cfg, cl, rn, pd, now)
    assert done == 1000
    assert (out.reason == abi.REASON_NONE).all()
    assert (out.start_time == now).all()
    # MinCpuTimeRatioFirst spreads load: first 128 jobs land on 128 distinct nodes
    assert len(set(out.alloc_node[:128].tolist())) == 128


@pytest.mark.parametrize("name", ["random_7", "random_8", "config2_small"])
def test_golden(oracle, name):
    """Oracle output is pinned by committed fixtures (tests/golden/make_golden.py)."""
    from tests.golden.make_golden import CASES
    case = CASES[name]()
    out, _, _ = oracle.node_select(*case[:4], case[4])
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    for f in abi.Placements.__dataclass_fields__:
        a = getattr(out, f)
        b = g[f]
        if a.dtype.names:
            assert a.tobytes() == b.tobytes(), f
        else:
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f


def test_feasible_micro(oracle):
    """GetFeasibleResourceInNode corner cases through the ABI structs."""
    cl = synth.make_cluster([(1, synth.node_row(8, 64 << 30, {0: 3, 1: 2}))], gres_entry_name=(0, 0))
    avail = cl.res_total[0].copy()
    req = np.zeros((), abi.RES_VIEW)
    req["cpu_raw"] = 2 * 256
    req["gres_total"][0] = 4
    req["gres_spec"][0] = 1
    ok, alloc = oracle.feasible(cl, req, avail)
    assert ok and alloc["core"][0] == 0b11 and alloc["gres"][0] == 0b111 and alloc["gres"][1] == 0b1
    req["gres_total"][0] = 6
    assert not oracle.feasible(cl, req, avail)[0]
    req["gres_total"][0] = 0
    req["gres_spec"][:] = 0
    req["cpu_raw"] = 384  # 1.5 cpus: no cores bound
    ok, alloc = oracle.feasible(cl, req, avail)
    assert ok and alloc["core"][0] == 0 and alloc["cpu_raw"] == 384
