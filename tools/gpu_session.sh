#!/bin/bash
# One gpurun call: GPU tests, a bench line, phase profile of the commit kernel.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputests.log 2>&1
tail -5 gpurun_out/gputests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -2 gpurun_out/bench.log


for c in 2 5; do timeout 300 python tools/profile_commit.py $c > gpurun_out/prof_config$c.log 2>&1; done
SEED_ID=3002 timeout 300 python tools/profile_commit.py 2 > gpurun_out/prof_config2_s3002.log 2>&1
head -3 gpurun_out/prof_config2.log gpurun_out/prof_config5.log gpurun_out/prof_config2_s3002.log
