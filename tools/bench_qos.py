"""Times the QoS post-filter behind NodeSelect on config 3
(the configuration BASELINE.json names for it): one tick + one filter pass.
Prints one JSON line. Usage: python tools/bench_qos.py [n_jobs n_nodes]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from cranesched_b200 import abi, synth
from cranesched_b200.scheduler import GpuScheduler

n_jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n_nodes = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000
cfg, cl, rn, pd, now = synth.config3(n_jobs=n_jobs, n_nodes=n_nodes)
table = synth.random_qos(3, cl, pd, tight=50.0)
s = GpuScheduler(cfg, 0)
s.set_cluster(cl)
out = s.node_select(now, rn, pd)           # warm-up
out = s.node_select(now, rn, pd)
t = s.timing()
started = int((out.reason == 0).sum())
reason = out.reason.copy()
t0 = time.perf_counter()
s.qos_filter(table.copy(), reason.copy())   # warm-up (first launch of the filter kernels) ...
out = s.node_select(now, rn, pd)            # ... and the tick's reasons back on the device
reason = out.reason.copy()
t0 = time.perf_counter()
s.qos_filter(table, reason)                 # H2D of the tables, kernels, D2H of reasons + usage
filt_ms = (time.perf_counter() - t0) * 1e3
qos_dev_ms = s.timing()["qos_ms"]
codes, counts = np.unique(reason, return_counts=True)
print(json.dumps({
    "workload": f"config3: {n_jobs} pending jobs x {n_nodes} nodes, 16 partitions, {table.n_qos} qos, {table.n_users} users, {table.n_accounts} accounts",
    "tick_ms": round(t["total_ms"], 3), "commit_ms": round(t["commit_ms"], 3),
    "decisions_per_s": round(n_jobs / (t["total_ms"] * 1e-3)),
    "started_now": started,
    "qos_filter_ms_host_to_host": round(filt_ms, 3),
    "qos_filter_ms_device": round(qos_dev_ms, 3),
    "qos_share_of_tick": round(qos_dev_ms / t["total_ms"], 4),
    "qos_decisions_per_s": round(started / (filt_ms * 1e-3)),
    "reasons_after_filter": {abi.REASON_STR.get(int(c), str(int(c))) or "started": int(n) for c, n in zip(codes, counts)},
}))
s.close()
