#!/bin/bash
# One gpurun call: GPU tests, phase profile of k_commit, sanitizer passes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputests.log 2>&1
timeout 300 python tools/profile_commit.py 2 > gpurun_out/prof_config2.log 2>&1
SEED_ID=3002 timeout 300 python tools/profile_commit.py 2 > gpurun_out/prof_config2_s3002.log 2>&1
timeout 300 python tools/profile_commit.py 5 > gpurun_out/prof_config5.log 2>&1
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=k_commit --print-limit 40 \
    python -m pytest tests/test_gpu_parity.py -q -x -k "random_sweep and (100 or 103 or 104 or 107)" > gpurun_out/sanitizer_$tool.log 2>&1
done
tail -3 gpurun_out/gputests.log
