#!/bin/bash
# A/B builds of the product library on the same GPU box. Variants are
# tools/variants/libcrane_sched_<V>.so (git-ignored); "B" is libcrane_sched.so.
#   tools/ab.sh            -> A B A B
#   VARIANTS="A B C" tools/ab.sh
for v in ${VARIANTS:-A B A B}; do
  if [ $v = B ]; then export CRANE_SCHED_LIB=$PWD/cranesched_b200/csrc/libcrane_sched.so; else export CRANE_SCHED_LIB=$PWD/tools/variants/libcrane_sched_$v.so; fi
  echo -n "$v "
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > /tmp/ab_$v.out 2> /tmp/ab_$v.err
  tail -1 /tmp/ab_$v.out | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['phases_ms']['commit_ms'])" 2>/dev/null || { echo "failed:"; tail -3 /tmp/ab_$v.err; }
done
