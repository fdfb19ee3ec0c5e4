#!/bin/bash
mkdir -p gpurun_out/r2b
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 900 -rA -k "register_xpass or config1_size or adjoint_vs_oracle or sums_leading or radix4 or step_host or full_gradient" 2>&1 | tail -80 > gpurun_out/r2b/pytest.log
tail -5 gpurun_out/r2b/pytest.log
for i in 1 2 3 4 5 6 7 8 9 10; do timeout 300 python -m pytest tests/test_gpu_gravity.py -q -m gpu -p no:cacheprovider -k "test_nbody_adjoint_vs_oracle and atomic" 2>&1 | tail -1; done | tee gpurun_out/r2b/adjoint_repeat.txt
