#!/bin/bash
# 2 GPUs: tiled deposit on slabs (single-GPU slab-descriptor test, world-2 parity, bench N=2 with / without)
mkdir -p gpurun_out/r2u
timeout 600 python -m pytest tests/test_gpu_sweep.py -q -m gpu -p no:cacheprovider --timeout 300 -k "slab_descriptor" 2>&1 | tail -6 | tee gpurun_out/r2u/pytest_sweep_slab.log
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu -p no:cacheprovider -k "copy_engine and 2" 2>&1 | tail -6 | tee gpurun_out/r2u/pytest_dist.log
for SW in 1 0; do
PMWD_SLAB_SWEEP=$SW NCCL_DEBUG=VERSION timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$SW \
   bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2u/bench_n2_sw$SW.json 2> gpurun_out/r2u/bench_n2_sw$SW.err
echo "bench n2 slab_sweep=$SW rc=$?"; grep -i "error\|Traceback" -A6 gpurun_out/r2u/bench_n2_sw$SW.err | head -20
done
python tools/bench_show.py gpurun_out/r2u/bench_n2_sw1.json gpurun_out/r2u/bench_n2_sw0.json
ls gpurun_out/*.log 2>/dev/null
