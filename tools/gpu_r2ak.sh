#!/bin/bash
# 1 GPU: the bench line of the final code (with the predicted-position re-sort), as the driver runs it
mkdir -p gpurun_out/r2ak
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2ak/bench_n1.json 2> gpurun_out/r2ak/bench_n1.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2ak/bench_n1.err
python tools/bench_show.py gpurun_out/r2ak/bench_n1.json
