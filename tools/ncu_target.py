import sys; sys.path.insert(0,'/root/repo')
from cranesched_b200 import synth
from cranesched_b200.scheduler import GpuScheduler
cfg, cl, rn, pd, now = synth.config2()
s = GpuScheduler(cfg, 0); s.set_cluster(cl)
out = s.node_select(now, rn, pd)
print(s.timing())
