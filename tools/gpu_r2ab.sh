#!/bin/bash
# 1 GPU: padded-spectrum lab; reciprocal-cell CIC kernels + 24-bit sort (tests, bench); deterministic-mode throughput
mkdir -p gpurun_out/r2ab
python tools/lab/fft2d_pad.py 1024 0 520 528 544 576 2>&1 | grep -v Warning | tee gpurun_out/r2ab/fft2d_pad.txt
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_sweep.py tests/test_gpu_cic.py -q -m gpu -p no:cacheprovider --timeout 300 2>&1 | tail -4 | tee gpurun_out/r2ab/pytest.log
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-context --e2e-steps 2 > gpurun_out/r2ab/bench_n1.json 2> gpurun_out/r2ab/bench_n1.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2ab/bench_n1.err
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-context --e2e-steps 1 --scatter-mode deterministic > gpurun_out/r2ab/bench_n1_det.json 2> gpurun_out/r2ab/bench_n1_det.err
echo "bench det rc=$?"; tail -c 300 gpurun_out/r2ab/bench_n1_det.err
python tools/bench_show.py gpurun_out/r2ab/bench_n1.json gpurun_out/r2ab/bench_n1_det.json 2>&1 | grep -E "=====| F | A "
