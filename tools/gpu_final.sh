#!/bin/bash
# 1 GPU, end of round 2: the whole GPU suite, smoke(), the bench lines as the driver runs them (+ reference arm),
# the ncu launch list of the same command and ncu --set full captures of the hand-written particle kernels
mkdir -p gpurun_out/final
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 900 2>&1 | tail -12 | tee gpurun_out/final/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/final/smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final/bench_n1.json 2> gpurun_out/final/bench_n1.err
echo "bench rc=$?"; tail -c 300 gpurun_out/final/bench_n1.err
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/final/bench_reference_arm.json 2> gpurun_out/final/bench_reference_arm.err
echo "reference arm rc=$?"
BENCH="python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-context --e2e-steps 1"
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    -k regex:'pmwd|fft|Radix|radix' -c 4000 --csv --log-file gpurun_out/final/launches_bench.csv $BENCH \
    > gpurun_out/final/launches_bench.log 2>&1
echo "launch list rc=$?"
for K in scatter_sweep_kernel gather3_kernel; do
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k regex:$K -s 20 -c 1 -o gpurun_out/final/prof_$K -f $BENCH --no-adjoint \
      > gpurun_out/final/prof_$K.log 2>&1
done
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:force_adj_gather_kernel -s 10 -c 1 -o gpurun_out/final/prof_force_adj_gather_kernel -f $BENCH \
    > gpurun_out/final/prof_force_adj_gather_kernel.log 2>&1
ls -la gpurun_out/final | tail -20
python tools/bench_show.py gpurun_out/final/bench_n1.json
