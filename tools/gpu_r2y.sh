#!/bin/bash
# 2 GPUs: two-source re-sort on slabs: world-2 parity + bench N=2 (forward + adjoint)
mkdir -p gpurun_out/r2y
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu -p no:cacheprovider -k "copy_engine and 2" 2>&1 | tail -6 | tee gpurun_out/r2y/pytest_dist.log
NCCL_DEBUG=VERSION timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
   bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2y/bench_n2.json 2> gpurun_out/r2y/bench_n2.err
echo "bench n2 rc=$?"; grep -i "error\|Traceback" -A6 gpurun_out/r2y/bench_n2.err | head -20
python tools/bench_show.py gpurun_out/r2y/bench_n2.json
cat gpurun_out/dist_worker_fail* 2>/dev/null | grep -v Warn | tail -30
