#!/bin/bash
# 1 GPU (last seconds of the budget): the full-size config-2 / config-3 tests on the final code
mkdir -p gpurun_out/r2al
timeout 110 python -m pytest tests/test_gpu_gravity.py -q -m gpu -p no:cacheprovider -x -k "config2 or config3 or full_size" --durations=5 2>&1 | tail -12 | tee gpurun_out/r2al/pytest_fullsize.log
