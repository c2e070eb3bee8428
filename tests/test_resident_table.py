"""The pending table kept resident on the device across ticks
(crane_sched_pending_append / _erase / set_running): rows
arrive in several submits, some are erased (started / cancelled), and the tick
must equal the oracle's NodeSelect over the surviving rows — through the
kernel-emulation library here, on the GPU in the `gpu`-marked twin."""
import numpy as np
import pytest

from cranesched_b200 import abi, synth
from cranesched_b200.scheduler import GpuScheduler


def _resident_tick(lib, case, seed):
    cfg, cl, rn, pd, now = case
    rng = np.random.Generator(np.random.PCG64(seed))
    s = GpuScheduler(cfg, 0, lib)
    try:
        s.set_cluster(cl)
        cuts = sorted(rng.choice(np.arange(1, pd.n), 2, replace=False).tolist())
        first = [s.pending_append(pd.take(np.arange(a, b))) for a, b in zip([0] + cuts, cuts + [pd.n])]
        assert first == [0] + cuts and s.pending_rows() == pd.n
        dead = np.sort(rng.choice(pd.n, pd.n // 5, replace=False))
        s.pending_erase(dead[: len(dead) // 2])
        s.pending_erase(dead[len(dead) // 2:])
        s.set_running(rn)
        s.run(now)
        out = s.fetch(abi.Placements.for_pending(pd))
        # a second tick on the same resident table gives the same answer
        s.set_running(rn)
        s.run(now)
        again = s.fetch(abi.Placements.for_pending(pd))
    finally:
        s.close()
    return out, again, dead


def _check(oracle, out, again, dead, case):
    cfg, cl, rn, pd, now = case
    assert not out.diff(again)
    alive = np.setdiff1d(np.arange(pd.n), dead)
    sub = pd.take(alive)
    ref, _, _ = oracle.node_select(cfg, cl, rn, sub, now)
    assert (out.reason[dead] == 255).all() and (out.n_alloc[dead] == 0).all()
    for f in ("reason", "start_time", "end_time", "n_alloc"):
        assert np.array_equal(getattr(ref, f), getattr(out, f)[alive]), f
    assert np.array_equal(ref.priority.view(np.uint64), out.priority[alive].view(np.uint64))
    rep = np.repeat(np.arange(pd.n), pd.node_num)
    rows = np.isin(rep, alive)
    assert np.array_equal(ref.alloc_node, out.alloc_node[rows])
    assert np.array_equal(ref.alloc_ntasks, out.alloc_ntasks[rows])
    assert ref.alloc_res.tobytes() == out.alloc_res[rows].tobytes()


@pytest.mark.parametrize("seed,kw", [(61, dict(fifo=False)), (62, dict(fifo=True, limit=60))])
def test_resident_table_emulated(oracle, emu_lib, seed, kw):
    case = synth.random_case(seed, n_jobs=80, n_nodes=24, n_parts=2, n_running=10, **kw)
    out, again, dead = _resident_tick(emu_lib, case, seed)
    _check(oracle, out, again, dead, case)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(70, 76))
def test_resident_table_gpu(oracle, gpu_lib, seed):
    case = synth.random_case(seed, n_jobs=500, n_nodes=60, n_parts=1 + seed % 3, n_running=30, fifo=bool(seed % 3 == 0),
                             limit=300 if seed % 2 else None)
    out, again, dead = _resident_tick(gpu_lib, case, seed)
    _check(oracle, out, again, dead, case)
