#!/bin/bash
mkdir -p gpurun_out/r2c
NCCL_DEBUG=INFO python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c/bench_n1.json 2> gpurun_out/r2c/bench_n1.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2c/bench_n1.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c/ref_n1.json 2> gpurun_out/r2c/ref_n1.err
echo "ref rc=$?"; tail -c 300 gpurun_out/r2c/ref_n1.err
