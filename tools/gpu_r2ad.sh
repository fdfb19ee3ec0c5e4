#!/bin/bash
# 1 GPU: templated sort-keys / permute kernels + streaming hints for the particle arrays in the gathers (A/B)
mkdir -p gpurun_out/r2ad
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_sweep.py tests/test_gpu_cic.py -q -m gpu -p no:cacheprovider --timeout 300 2>&1 | tail -4 | tee gpurun_out/r2ad/pytest.log
for CS in 1 0; do
PMWD_PTCL_CS=$CS python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-context --e2e-steps 1 > gpurun_out/r2ad/bench_n1_cs$CS.json 2> gpurun_out/r2ad/bench_n1_cs$CS.err
echo "bench cs=$CS rc=$?"; tail -c 300 gpurun_out/r2ad/bench_n1_cs$CS.err
done
python tools/bench_show.py gpurun_out/r2ad/bench_n1_cs1.json gpurun_out/r2ad/bench_n1_cs0.json 2>&1 | grep -E "=====|gather|other|scatter"
