#!/bin/bash
# round 2, first GPU call: the whole GPU suite without -x incl. everything gated, + pending timings
mkdir -p gpurun_out/r2a
export PMWD_RUN_UNVALIDATED=1
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -rA 2>&1 | tail -200 > gpurun_out/r2a/pytest_full.log
tail -5 gpurun_out/r2a/pytest_full.log
echo "--- radix-4 2048" ; timeout 300 python tools/time_xpass.py 2048 256 2>&1 | tail -1 | tee gpurun_out/r2a/xpass2048_r4.txt
echo "--- cluster 2048" ; PMWD_XPASS16_2048=1 timeout 300 python tools/time_xpass.py 2048 256 2>&1 | tail -1 | tee gpurun_out/r2a/xpass2048_cluster.txt
echo "--- torch baseline"; timeout 300 python tools/torch_baseline.py 256 5 2>&1 | tail -1 | tee gpurun_out/r2a/torch_baseline.txt
