#!/bin/bash
# 1 GPU: sweep kernel round 2 (parity fix for mixed-parity chunks, 2 chunks in flight, batched straggler slots)
mkdir -p gpurun_out/r2s
timeout 900 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_fused.py -q -m gpu -p no:cacheprovider --timeout 300 2>&1 | tail -8 | tee gpurun_out/r2s/pytest_sweep.log
timeout 600 python -m pytest tests/test_gpu_gravity.py -q -m gpu -p no:cacheprovider --timeout 300 -k "nbody_step_host" 2>&1 | tail -8 | tee gpurun_out/r2s/pytest_host.log
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-context > gpurun_out/r2s/bench_n1.json 2> gpurun_out/r2s/bench_n1.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2s/bench_n1.err
python tools/time_host_step.py 512 3 > gpurun_out/r2s/host_trace.txt 2>&1; tail -30 gpurun_out/r2s/host_trace.txt
BENCH="python bench.py --no-cpu-baseline --e2e-steps 1 --no-adjoint --no-context"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:scatter_sweep_kernel -s 45 -c 1 -o gpurun_out/r2s/prof_sweep_s45 -f $BENCH --steps 47 \
    > gpurun_out/r2s/prof_sweep_s45.log 2>&1
python tools/bench_show.py gpurun_out/r2s/bench_n1.json
