#!/bin/bash
mkdir -p gpurun_out/r2m
python tools/lab/fft2d_variants.py 1024 2>&1 | tail -1 | tee gpurun_out/r2m/fft2d_variants.txt
for R in 3 4 5 6; do
python bench.py --gpus 1 --no-cpu-baseline --no-context --e2e-steps 1 --reorder-every $R > gpurun_out/r2m/bench_reorder$R.json 2> gpurun_out/r2m/bench_reorder$R.err
echo "reorder_every=$R rc=$?"
done
