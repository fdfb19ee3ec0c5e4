#!/bin/bash
mkdir -p gpurun_out/r2d
timeout 900 python -m pytest tests/test_gpu_sweep.py -q -m gpu -p no:cacheprovider --timeout 300 -rA -x 2>&1 | tail -60 > gpurun_out/r2d/pytest_sweep.log
tail -15 gpurun_out/r2d/pytest_sweep.log
