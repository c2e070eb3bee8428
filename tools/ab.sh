#!/bin/bash
# A/B two builds of the product library on the same GPU box:
#   A = cranesched_b200/csrc/libcrane_sched_A.so, B = libcrane_sched.so
for v in A B A B; do
  if [ $v = A ]; then export CRANE_SCHED_LIB=$PWD/cranesched_b200/csrc/libcrane_sched_A.so; else export CRANE_SCHED_LIB=$PWD/cranesched_b200/csrc/libcrane_sched.so; fi
  echo -n "$v "
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['phases_ms']['commit_ms'])"
done
