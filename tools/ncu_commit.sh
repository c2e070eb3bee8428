#!/bin/bash
# ncu capture of the commit kernel on config 2 (one launch), with source correlation.
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:k_commit -c 1 \
  -o gpurun_out/commit_v2 -f python tools/ncu_target.py > gpurun_out/ncu_commit.log 2>&1
tail -3 gpurun_out/ncu_commit.log
ls -la gpurun_out/*.ncu-rep
