"""One-box comparison of library variants (tools/variants/libcrane_sched_<V>.so; B = the in-tree build) on
config 2 (default draw and the over-subscribed seed 3002) and config 5: commit_ms of the second tick.
  python tools/ab_multi.py A B"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cranesched_b200 import synth
from cranesched_b200.scheduler import GpuScheduler, LIB_PATH
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cases = [("config2", synth.config2()), ("config2_s3002", synth.config2(seed_id=3002)), ("config5", synth.config5())]
for v in sys.argv[1:] or ["A", "B"]:
    lib = LIB_PATH if v == "B" else os.path.join(root, "tools", "variants", "libcrane_sched_%s.so" % v)
    row = []
    for name, (cfg, cl, rn, pd, now) in cases:
        s = GpuScheduler(cfg, 0, lib); s.set_cluster(cl)
        s.node_select(now, rn, pd); s.node_select(now, rn, pd)
        row.append("%s %.1f" % (name, s.timing()["commit_ms"])); s.close()
    print(v, " | ".join(row), flush=True)
