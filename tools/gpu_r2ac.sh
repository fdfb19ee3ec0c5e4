#!/bin/bash
# 1 GPU: deterministic tiled deposit: unit tests, the deterministic cases of the gravity suite, bench in deterministic mode
mkdir -p gpurun_out/r2ac
timeout 600 python -m pytest tests/test_gpu_sweep.py -q -m gpu -p no:cacheprovider --timeout 300 2>&1 | tail -8 | tee gpurun_out/r2ac/pytest_sweep.log
timeout 900 python -m pytest tests/test_gpu_gravity.py -q -m gpu -p no:cacheprovider --timeout 600 -k "deterministic or reproduc" 2>&1 | tail -8 | tee gpurun_out/r2ac/pytest_det.log
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-context --e2e-steps 1 --scatter-mode deterministic > gpurun_out/r2ac/bench_n1_det.json 2> gpurun_out/r2ac/bench_n1_det.err
echo "bench det rc=$?"; tail -c 400 gpurun_out/r2ac/bench_n1_det.err
python tools/bench_show.py gpurun_out/r2ac/bench_n1_det.json 2>&1 | grep -E "=====| F | A "
