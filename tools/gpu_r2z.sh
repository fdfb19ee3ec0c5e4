#!/bin/bash
# 2 GPUs: slab host-step API (parity inside the dist worker) + bench N=2 (forward, adjoint with warm-up, e2e through it)
mkdir -p gpurun_out/r2z
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu -p no:cacheprovider -k "2" 2>&1 | tail -6 | tee gpurun_out/r2z/pytest_dist.log
NCCL_DEBUG=VERSION timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
   bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2z/bench_n2.json 2> gpurun_out/r2z/bench_n2.err
echo "bench n2 rc=$?"; grep -i "error\|Traceback" -A6 gpurun_out/r2z/bench_n2.err | head -20
python tools/bench_show.py gpurun_out/r2z/bench_n2.json
cat gpurun_out/dist_worker_fail* 2>/dev/null | grep -v Warn | tail -30
