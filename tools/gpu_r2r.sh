#!/bin/bash
# 8 GPUs: the world-8 slab parity worker with Eulerian ownership, full log (it failed inside the pytest wrapper);
# control without migration only if it fails
mkdir -p gpurun_out/r2r
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 \
  tests/dist_gpu_worker.py > gpurun_out/r2r/worker8.log 2>&1
rc=$?
echo "worker rc=$rc"
grep -v "Warning\|warn" gpurun_out/r2r/worker8.log | grep -i "error\|assert\|ok:" | head -30
if [ $rc -ne 0 ]; then
PMWD_MIGRATE=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 \
  tests/dist_gpu_worker.py > gpurun_out/r2r/worker8_nomig.log 2>&1
echo "control (PMWD_MIGRATE=0) rc=$?"
grep -v "Warning\|warn" gpurun_out/r2r/worker8_nomig.log | grep -i "error\|assert\|ok:" | head -30
fi
