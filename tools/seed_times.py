"""Tick time of config 2 for the per-rank seeds of the weak-scaling bench."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cranesched_b200 import synth, sharding
from cranesched_b200.scheduler import GpuScheduler
for rank in range(8):
    cfg, cl, rn, pd, now = sharding.shard_workload(2, rank, 8)  # seed 1000*rank + 2
    s = GpuScheduler(cfg, 0); s.set_cluster(cl)
    out = s.node_select(now, rn, pd); out = s.node_select(now, rn, pd)
    t = s.timing(); s.close()
    jobs = np.bincount(pd.partition, minlength=4)
    print(rank, "commit_ms %.1f" % t["commit_ms"], "jobs/part", jobs.tolist(), "started", int((out.reason == 0).sum()),
          "reserved", int(((out.reason != 0) & (out.n_alloc > 0)).sum()), "unplaced", int((out.n_alloc == 0).sum()), flush=True)
