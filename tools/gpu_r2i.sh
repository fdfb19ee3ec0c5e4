#!/bin/bash
# 2 GPUs: pipelined slab FFT parity + bench A/B; fused LPT kernels parity (1 GPU)
mkdir -p gpurun_out/r2i
timeout 600 python -m pytest tests/test_gpu_gravity.py -q -m gpu -p no:cacheprovider -k "lpt or full_gradient" 2>&1 | tail -5 | tee gpurun_out/r2i/pytest_lpt.log
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -15 | tee gpurun_out/r2i/pytest_dist.log
for PIPE in 2 4 1; do
PMWD_PIPE=$PIPE timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 20 --warmup 5 --e2e-steps 1 > gpurun_out/r2i/bench_n2_pipe$PIPE.json 2> gpurun_out/r2i/bench_n2_pipe$PIPE.err
echo "bench n2 pipe=$PIPE rc=$?"; tail -c 200 gpurun_out/r2i/bench_n2_pipe$PIPE.err
done
