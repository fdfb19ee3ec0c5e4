#!/bin/bash
# 1 GPU: the whole 63-step run (bench.py defaults: W=3, K=60) on the final code, for the late-time (clustered) average
mkdir -p gpurun_out/r2ah
python bench.py --no-context > gpurun_out/r2ah/bench_n1_whole_run.json 2> gpurun_out/r2ah/bench_n1_whole_run.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2ah/bench_n1_whole_run.err
python tools/bench_show.py gpurun_out/r2ah/bench_n1_whole_run.json
