#!/bin/bash
# 1 GPU: validate the adjoint fusions + aligned sweep ring (full suite), then tile-shape A/B at full size
mkdir -p gpurun_out/r2k
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 900 -x 2>&1 | tail -12 > gpurun_out/r2k/pytest_full.log
tail -4 gpurun_out/r2k/pytest_full.log
for V in "16 32" "8 64" "8 32" "16 64"; do
set -- $V
PMWD_SWEEP_TY=$1 PMWD_SWEEP_BW=$2 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-context --e2e-steps 2 > gpurun_out/r2k/bench_ty$1_bw$2.json 2> gpurun_out/r2k/bench_ty$1_bw$2.err
echo "ty=$1 bw=$2 rc=$?"
done
