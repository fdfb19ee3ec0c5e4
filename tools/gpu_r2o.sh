#!/bin/bash
# 2 GPUs: Eulerian ownership (migration at every re-sort) parity + bench A/B
mkdir -p gpurun_out/r2o
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu -p no:cacheprovider -x -k "copy_engine" 2>&1 | tail -15 | tee gpurun_out/r2o/pytest_dist.log
for MIG in 1 0; do
PMWD_MIGRATE=$MIG timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --e2e-steps 1 > gpurun_out/r2o/bench_n2_mig$MIG.json 2> gpurun_out/r2o/bench_n2_mig$MIG.err
echo "bench n2 migrate=$MIG rc=$?"; grep -i "error\|Traceback" -A5 gpurun_out/r2o/bench_n2_mig$MIG.err | head -20
done
