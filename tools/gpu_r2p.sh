#!/bin/bash
# 8 GPUs: Eulerian ownership parity at world 8 + bench N=8 (driver arguments)
mkdir -p gpurun_out/r2p
timeout 400 python -m pytest tests/test_gpu_dist.py -q -m gpu -p no:cacheprovider -k "copy_engine and 8" 2>&1 | tail -6 | tee gpurun_out/r2p/pytest_dist.log
NCCL_DEBUG=VERSION timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
   bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2p/bench_n8.json 2> gpurun_out/r2p/bench_n8.err
echo "bench n8 rc=$?"; grep -i "error\|Traceback" -A6 gpurun_out/r2p/bench_n8.err | head -20
