#!/bin/bash
mkdir -p gpurun_out/r2l
for V in "16 32" "8 64"; do
set -- $V
PMWD_SWEEP_TY=$1 PMWD_SWEEP_BW=$2 python bench.py --gpus 1 --no-cpu-baseline --no-context --e2e-steps 1 > gpurun_out/r2l/bench_full_ty$1_bw$2.json 2> gpurun_out/r2l/bench_full_ty$1_bw$2.err
echo "ty=$1 bw=$2 rc=$?"
done
