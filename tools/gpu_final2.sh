#!/bin/bash
# 1 GPU: last full pass on the final code: whole GPU suite, smoke(), bench as the driver runs it, launch list
mkdir -p gpurun_out/final2
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 900 2>&1 | tail -6 | tee gpurun_out/final2/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/final2/smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final2/bench_n1.json 2> gpurun_out/final2/bench_n1.err
echo "bench rc=$?"; tail -c 300 gpurun_out/final2/bench_n1.err
BENCH="python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-context --e2e-steps 1"
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    -k regex:'pmwd|fft|Radix|radix' -c 4000 --csv --log-file gpurun_out/final2/launches_bench.csv $BENCH \
    > gpurun_out/final2/launches_bench.log 2>&1
echo "launch list rc=$?"
python tools/bench_show.py gpurun_out/final2/bench_n1.json
