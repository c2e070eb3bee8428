#!/bin/bash
# A/B of library variants on config 2, default seed and seed 3002 (commit ms).
for v in ${VARIANTS:-B C B C}; do
  if [ $v = B ]; then export CRANE_SCHED_LIB=$PWD/cranesched_b200/csrc/libcrane_sched.so; else export CRANE_SCHED_LIB=$PWD/cranesched_b200/csrc/libcrane_sched_$v.so; fi
  python - <<'P'
import os,sys
sys.path.insert(0,os.getcwd())
from cranesched_b200 import synth
from cranesched_b200.scheduler import GpuScheduler
out=[]
for seed in (2,3002):
    cfg,cl,rn,pd,now=synth.config2(seed_id=seed)
    s=GpuScheduler(cfg,0); s.set_cluster(cl)
    s.node_select(now,rn,pd); s.node_select(now,rn,pd)
    out.append(round(s.timing()["commit_ms"],1)); s.close()
print(os.environ["CRANE_SCHED_LIB"][-12:], out)
P
done
