#!/bin/bash
# 8 GPUs: slab parity at world 4 and 8, bench N=8 (config 4 + config 5: forward + adjoint)
mkdir -p gpurun_out/r2h
timeout 500 python -m pytest tests/test_gpu_dist.py -q -m gpu -p no:cacheprovider -k "copy_engine and (4 or 8)" 2>&1 | tail -8 | tee gpurun_out/r2h/pytest_dist.log
NCCL_DEBUG=VERSION timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
   bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2h/bench_n8.json 2> gpurun_out/r2h/bench_n8.err
echo "bench n8 rc=$?"; tail -c 400 gpurun_out/r2h/bench_n8.err
