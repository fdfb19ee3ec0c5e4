#!/bin/bash
# 8 GPUs: final multi-GPU state: world-8 parity (Eulerian ownership + tiled deposit on slabs + two-source re-sort +
# slab host-step) and bench N=8 as the driver runs it
mkdir -p gpurun_out/r2aa
timeout 400 python -m pytest tests/test_gpu_dist.py -q -m gpu -p no:cacheprovider -k "copy_engine and 8" 2>&1 | tail -6 | tee gpurun_out/r2aa/pytest_dist.log
NCCL_DEBUG=VERSION timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 \
   bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2aa/bench_n8.json 2> gpurun_out/r2aa/bench_n8.err
echo "bench n8 rc=$?"; grep -i "error\|Traceback" -A6 gpurun_out/r2aa/bench_n8.err | head -20
python tools/bench_show.py gpurun_out/r2aa/bench_n8.json
cp gpurun_out/dist_worker_fail* gpurun_out/r2aa/ 2>/dev/null; cat gpurun_out/dist_worker_fail* 2>/dev/null | grep -v Warn | grep -i "rank.*error" | head
