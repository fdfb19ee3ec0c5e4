#!/bin/bash
# 1 GPU: re-sort by predicted positions (order centred on the forces until the next re-sort): tests + bench on / off
mkdir -p gpurun_out/r2ai
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_sweep.py -q -m gpu -p no:cacheprovider --timeout 300 2>&1 | tail -3 | tee gpurun_out/r2ai/pytest_a.log
timeout 900 python -m pytest tests/test_gpu_gravity.py -q -m gpu -p no:cacheprovider --timeout 600 -k "nbody_config1 or adjoint_vs_oracle or reproduc or full_gradient" 2>&1 | tail -3 | tee gpurun_out/r2ai/pytest_b.log
for P in 1 0; do
PMWD_SORT_PREDICT=$P python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-context --e2e-steps 1 > gpurun_out/r2ai/bench_pred$P.json 2> gpurun_out/r2ai/bench_pred$P.err
echo "predict=$P rc=$?"; tail -c 200 gpurun_out/r2ai/bench_pred$P.err
done
python tools/bench_show.py gpurun_out/r2ai/bench_pred1.json gpurun_out/r2ai/bench_pred0.json 2>&1 | grep -E "=====|scatter|gather|other"
