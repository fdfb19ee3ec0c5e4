#!/bin/bash
# 2 GPUs: slab parity (copy-engine and store-kernel transposes) + bench A/B incl. the adjoint leg
mkdir -p gpurun_out/r2g
timeout 1200 python -m pytest tests/test_gpu_dist.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -15 | tee gpurun_out/r2g/pytest_dist.log
for CE in 1 0; do
PMWD_P2P_CE=$CE NCCL_DEBUG=VERSION timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2g/bench_n2_ce$CE.json 2> gpurun_out/r2g/bench_n2_ce$CE.err
echo "bench n2 ce=$CE rc=$?"; tail -c 300 gpurun_out/r2g/bench_n2_ce$CE.err
done
OMP_NUM_THREADS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2g/ref_n2.json 2> gpurun_out/r2g/ref_n2.err
echo "ref rc=$?"
