#!/bin/bash
# 2 GPUs: the predicted-position re-sort on slabs (two-source keys with the arrivals' velocities): world-2 parity
mkdir -p gpurun_out/r2aj
timeout 300 python -m pytest tests/test_gpu_dist.py -q -m gpu -p no:cacheprovider -k "copy_engine and 2" 2>&1 | tail -5 | tee gpurun_out/r2aj/pytest_dist.log
cat gpurun_out/dist_worker_fail* 2>/dev/null | grep -v Warn | grep -i "rank.*error" | head
