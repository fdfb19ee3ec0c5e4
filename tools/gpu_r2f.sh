#!/bin/bash
mkdir -p gpurun_out/r2f
timeout 900 python -m pytest tests/test_gpu_sweep.py -q -m gpu -p no:cacheprovider --timeout 300 2>&1 | tail -8
python bench.py --gpus 1 --no-cpu-baseline --no-context --e2e-steps 1 > gpurun_out/r2f/bench_n1_full.json 2> gpurun_out/r2f/bench_n1_full.err
echo "bench full rc=$?"; tail -c 300 gpurun_out/r2f/bench_n1_full.err
BENCH="python bench.py --no-cpu-baseline --e2e-steps 1 --no-adjoint --no-context"
for S in 8 45; do
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:scatter_sweep_kernel -s $S -c 1 -o gpurun_out/r2f/prof_sweep_s$S -f $BENCH --steps 47 \
    > gpurun_out/r2f/prof_sweep_s$S.log 2>&1
done
