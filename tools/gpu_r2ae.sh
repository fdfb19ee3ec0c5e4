#!/bin/bash
# 4 GPUs: the missing point of the scaling table on the final multi-GPU code
mkdir -p gpurun_out/r2ae
NCCL_DEBUG=VERSION timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29591 \
   bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2ae/bench_n4.json 2> gpurun_out/r2ae/bench_n4.err
echo "bench n4 rc=$?"; grep -i "error\|Traceback" -A6 gpurun_out/r2ae/bench_n4.err | head -20
python tools/bench_show.py gpurun_out/r2ae/bench_n4.json
