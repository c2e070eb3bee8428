"""Debug helper: the random sweep of tests/test_gpu_parity.py against whatever
library CRANE_SCHED_LIB names (the test fixtures always load the in-tree build)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cranesched_b200 import synth
from oracle import pyoracle as o
from tests.helpers import run_sched

shapes = [dict(n_nodes=7, n_parts=1), dict(n_nodes=8, n_parts=1), dict(n_nodes=20, n_parts=2),
          dict(n_nodes=70, n_parts=1), dict(n_nodes=130, n_parts=3), dict(n_nodes=33, n_parts=4)]
bad = 0
cases = []
for seed in range(100, 124):
    kw = dict(shapes[seed % len(shapes)])
    kw.update(n_jobs=250 + 37 * (seed % 11), n_running=seed % 40, fifo=bool(seed % 5 == 0), short=bool(seed % 2))
    if seed % 7 == 0:
        kw["max_jobs_per_node"] = 9
    cases.append((seed, synth.random_case(seed, **kw)))
for seed in (40, 41, 42):
    cases.append((seed, synth.random_case(seed, n_jobs=600, n_nodes=12, n_parts=2, n_running=10, max_jobs_per_node=16)))
cases.append(("cfg2", synth.config2(n_jobs=6000, n_nodes=600)))
for seed, case in cases:
    ref, _, _ = o.node_select(*case[:4], case[4])
    got, _ = run_sched(case, None)
    d = ref.diff(got)
    if d:
        bad += 1
        print(seed, "DIFF", d[:2])
print("cases", len(cases), "bad", bad)
