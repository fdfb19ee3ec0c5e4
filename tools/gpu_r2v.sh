#!/bin/bash
# 2 GPUs: bench N=2 with migrate / resort phases timed (tiled deposit on slabs on / off)
mkdir -p gpurun_out/r2v
for SW in 1 0; do
PMWD_SLAB_SWEEP=$SW NCCL_DEBUG=VERSION timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$SW \
   bench.py --gpus 2 --steps 20 --warmup 5 --no-adjoint > gpurun_out/r2v/bench_n2_sw$SW.json 2> gpurun_out/r2v/bench_n2_sw$SW.err
echo "bench n2 slab_sweep=$SW rc=$?"; grep -i "error\|Traceback" -A6 gpurun_out/r2v/bench_n2_sw$SW.err | head -20
done
python tools/bench_show.py gpurun_out/r2v/bench_n2_sw1.json gpurun_out/r2v/bench_n2_sw0.json
