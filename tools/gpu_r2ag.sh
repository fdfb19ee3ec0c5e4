#!/bin/bash
# 1 GPU: planes per x segment of the deposit's work items: 128 instead of 64 (fewer planes that take REDs)
mkdir -p gpurun_out/r2ag
PMWD_SWEEP_LX=128 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-context --e2e-steps 1 > gpurun_out/r2ag/bench_lx128.json 2> gpurun_out/r2ag/bench_lx128.err
echo "lx=128 rc=$?"; tail -c 300 gpurun_out/r2ag/bench_lx128.err
python tools/bench_show.py gpurun_out/r2ag/bench_lx128.json 2>&1 | grep -E "=====|scatter|gather|other"
