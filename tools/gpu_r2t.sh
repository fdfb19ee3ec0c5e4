#!/bin/bash
# 1 GPU: L2-resident chunked 2-D FFTs (lab), sweep kernel round 3, host-step overlap fix
mkdir -p gpurun_out/r2t
python tools/lab/fft2d_chunks.py 1024 0 2 4 8 16 32 64 2>&1 | grep -v Warning | tee gpurun_out/r2t/fft2d_chunks.txt
timeout 600 python -m pytest tests/test_gpu_sweep.py -q -m gpu -p no:cacheprovider --timeout 300 2>&1 | tail -4 | tee gpurun_out/r2t/pytest_sweep.log
timeout 600 python -m pytest tests/test_gpu_gravity.py -q -m gpu -p no:cacheprovider --timeout 300 -k "nbody_step_host" 2>&1 | tail -4 | tee gpurun_out/r2t/pytest_host.log
for CH in 0 8; do
PMWD_FFT2D_CHUNK=$CH python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-context > gpurun_out/r2t/bench_n1_chunk$CH.json 2> gpurun_out/r2t/bench_n1_chunk$CH.err
echo "bench chunk=$CH rc=$?"; tail -c 300 gpurun_out/r2t/bench_n1_chunk$CH.err
done
python tools/bench_show.py gpurun_out/r2t/bench_n1_chunk0.json gpurun_out/r2t/bench_n1_chunk8.json
