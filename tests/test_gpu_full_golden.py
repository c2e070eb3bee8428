"""Whole-queue parity at the BASELINE.json sizes: every output column of the
GPU result (reason, fp64 priority bits, start/end, node sets, core and slot
masks) is compared with the CPU oracle's result of the same synthetic case
through SHA-256 digests precomputed by tests/golden/make_full_golden.py (the
oracle needs minutes to hours per case, single-threaded like the reference).
A case whose digest has not been generated yet is skipped, not passed."""
import json
import os

import numpy as np
import pytest

from cranesched_b200 import abi
from tests.golden.make_full_golden import CASES, DIGESTS, COLUMNS, digest_of
from tests.helpers import check_invariants, run_sched

pytestmark = pytest.mark.gpu


def _digests():
    if not os.path.exists(DIGESTS):
        return {}
    with open(DIGESTS) as f:
        return json.load(f)


@pytest.mark.parametrize("name", sorted(CASES))
def test_full_queue_digest(gpu_lib, name):
    want = _digests().get(name)
    if want is None:
        pytest.skip("no oracle digest for %s yet (tests/golden/make_full_golden.py)" % name)
    mk, keep = CASES[name]
    case = mk()
    assert case[3].n == want["n_jobs"] and case[1].n_nodes == want["n_nodes"]
    got, _ = run_sched(case, gpu_lib)
    have = digest_of(got)
    bad = [c for c in COLUMNS if have[c] != want[c]]
    if bad and keep:
        g = np.load(os.path.join(os.path.dirname(DIGESTS), name + ".npz"))
        ref = abi.Placements(**{f: g[f] for f in abi.Placements.__dataclass_fields__})
        raise AssertionError("columns differ: %s\n%s" % (bad, "\n".join(ref.diff(got)[:12])))
    assert not bad, "columns differ from the oracle's: %s (reason_hist got %s want %s)" % (
        bad, have["reason_hist"], want["reason_hist"])
    assert have["n_started"] == want["n_started"] and have["n_reserved"] == want["n_reserved"]
    if case[3].n <= 20_000:
        check_invariants(case, got)
