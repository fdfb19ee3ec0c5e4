#!/bin/bash
# 1 GPU: re-sort rewrite (two-source sort / permute, capacity buffers) under the single-GPU suite
mkdir -p gpurun_out/r2w
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_sweep.py -q -m gpu -p no:cacheprovider --timeout 300 2>&1 | tail -6 | tee gpurun_out/r2w/pytest_a.log
timeout 900 python -m pytest tests/test_gpu_gravity.py -q -m gpu -p no:cacheprovider --timeout 600 -k "nbody or adjoint or reorder or store" 2>&1 | tail -6 | tee gpurun_out/r2w/pytest_b.log
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-context > gpurun_out/r2w/bench_n1.json 2> gpurun_out/r2w/bench_n1.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2w/bench_n1.err
python tools/bench_show.py gpurun_out/r2w/bench_n1.json | head -3
