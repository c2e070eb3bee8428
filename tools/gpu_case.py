"""Debug helper: run one random_case several times on the GPU and print the diff against the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cranesched_b200 import synth
from oracle import pyoracle as o
from tests.helpers import run_sched

seed = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
case = synth.random_case(seed, n_jobs=600, n_nodes=12, n_parts=2, n_running=10, max_jobs_per_node=16)
ref, _, _ = o.node_select(*case[:4], case[4])
pd = case[3]
for r in range(reps):
    got, _ = run_sched(case, None)
    bad = [i for i in range(pd.n) if ref.reason[i] != got.reason[i] or ref.n_alloc[i] != got.n_alloc[i] or ref.start_time[i] != got.start_time[i]
           or (ref.alloc_node[ref.alloc_off[i]:ref.alloc_off[i] + ref.n_alloc[i]] != got.alloc_node[got.alloc_off[i]:got.alloc_off[i] + got.n_alloc[i]]).any()]
    print("rep", r, "bad jobs", len(bad), bad[:8])
    order = np.argsort(-ref.priority, kind="stable")
    rank = np.empty(pd.n, int); rank[order] = np.arange(pd.n)
    for i in sorted(bad, key=lambda i: rank[i])[:5]:
        print("  job", i, "rank", rank[i], "part", pd.partition[i], "K", pd.node_num[i], "excl", pd.exclusive[i],
              "ref", ref.reason[i], ref.alloc_node[ref.alloc_off[i]:ref.alloc_off[i] + ref.n_alloc[i]], ref.start_time[i] - case[4],
              "got", got.reason[i], got.alloc_node[got.alloc_off[i]:got.alloc_off[i] + got.n_alloc[i]], got.start_time[i] - case[4])
