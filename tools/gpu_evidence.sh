#!/bin/bash
# Sanitizer passes over k_commit2 (small cases), launch list + full ncu capture of the bench command.
mkdir -p gpurun_out
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --kernel-name kns=k_commit2 --print-limit 30 \
    python -m pytest tests/test_gpu_parity.py -q -x -k "(random_sweep and (100 or 103 or 104 or 107)) or (general_task and 201) or (bestfit and 301) or (overlapping and 601) or (test_reservations and 501)" > gpurun_out/sanitizer_$tool.log 2>&1
  tail -3 gpurun_out/sanitizer_$tool.log
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:k_commit2 -c 1 \
  -o gpurun_out/commit_v2 -f python tools/ncu_target.py > gpurun_out/ncu_commit.log 2>&1
ls -la gpurun_out | tail -8
