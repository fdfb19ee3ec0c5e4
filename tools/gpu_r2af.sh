#!/bin/bash
# 1 GPU: tile shape of the deposit again with the final kernel: 8 x 64 (default) vs 16 x 32 (same ring size, half the
# y-crossings that make up ~90 % of the stragglers)
mkdir -p gpurun_out/r2af
for V in "16 32" "8 64"; do
set -- $V
PMWD_SWEEP_TY=$1 PMWD_SWEEP_BW=$2 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-context --e2e-steps 1 > gpurun_out/r2af/bench_ty$1_bw$2.json 2> gpurun_out/r2af/bench_ty$1_bw$2.err
echo "ty=$1 bw=$2 rc=$?"
done
python tools/bench_show.py gpurun_out/r2af/bench_ty16_bw32.json gpurun_out/r2af/bench_ty8_bw64.json 2>&1 | grep -E "=====|scatter|gather|other"
