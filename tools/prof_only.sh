mkdir -p gpurun_out
SEED_ID=3002 timeout 300 python tools/profile_commit.py 2 > gpurun_out/prof_config2_s3002.log 2>&1
timeout 300 python tools/profile_commit.py 5 > gpurun_out/prof_config5.log 2>&1
grep -A20 "partition 2" gpurun_out/prof_config2_s3002.log; head -30 gpurun_out/prof_config5.log
