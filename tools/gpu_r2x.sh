#!/bin/bash
# 1 GPU: is the adjoint slowdown of r2w (88 vs 78 ms/step) reproducible?  two runs
mkdir -p gpurun_out/r2x
for R in 1 2; do
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-context --e2e-steps 1 > gpurun_out/r2x/bench_n1_$R.json 2> gpurun_out/r2x/bench_n1_$R.err
echo "bench rc=$?"
done
python tools/bench_show.py gpurun_out/r2x/bench_n1_1.json gpurun_out/r2x/bench_n1_2.json | grep -E "=====|other"
