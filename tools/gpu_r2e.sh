#!/bin/bash
# sweep scatter at full size: bench (driver args), full GPU suite, ncu of the sweep kernel early + late
mkdir -p gpurun_out/r2e
python bench.py --gpus 1 --steps 20 --warmup 5 --no-context > gpurun_out/r2e/bench_n1_s20.json 2> gpurun_out/r2e/bench_n1_s20.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r2e/bench_n1_s20.err
python bench.py --gpus 1 --no-cpu-baseline --no-context --e2e-steps 1 > gpurun_out/r2e/bench_n1_full.json 2> gpurun_out/r2e/bench_n1_full.err
echo "bench full rc=$?"
PMWD_SWEEP=0 python bench.py --gpus 1 --no-cpu-baseline --no-context --e2e-steps 1 > gpurun_out/r2e/bench_n1_full_nosweep.json 2> gpurun_out/r2e/bench_n1_full_nosweep.err
echo "bench nosweep rc=$?"
BENCH="python bench.py --no-cpu-baseline --e2e-steps 1 --no-adjoint --no-context"
for S in 8 45; do
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:scatter_sweep_kernel -s $S -c 1 -o gpurun_out/r2e/prof_sweep_s$S -f $BENCH --steps 47 \
    > gpurun_out/r2e/prof_sweep_s$S.log 2>&1
done
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 900 2>&1 | tail -15 > gpurun_out/r2e/pytest_full.log
tail -5 gpurun_out/r2e/pytest_full.log
