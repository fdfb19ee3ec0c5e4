#!/bin/bash
# ncu evidence for bench.py's kernels (run under gpurun, 1 GPU).  Outputs -> gpurun_out/<tag>/.
# usage: tools/profile.sh <tag>          (round 2: this is what tools/gpu_final.sh ran for profiles/r02_final_*)
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
BENCH="python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-context --e2e-steps 1"
# every launch of our kernels + cuFFT + cub with its device time (cold-cache, serialised: compare SHARES
# with kernels.*.share_of_step / adjoint.kernels.* of the bench line; tools/summarize_launches.py)
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    -k regex:'pmwd|fft|Radix|radix' -c 4000 --csv --log-file $OUT/launches_bench.csv $BENCH \
    > $OUT/launches_bench.log 2>&1
# full captures of the hand-written particle kernels inside the run (-s skips earlier launches; the adjoint
# gather needs the adjoint leg).  The x-pass kernels are data independent: capture them alone,
# `ncu --set full ... python tools/time_xpass.py 1024`.  tools/summarize_ncu.py extracts the roofline metrics.
for K in scatter_sweep_kernel gather3_kernel; do
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k regex:$K -s 20 -c 1 -o $OUT/prof_$K -f $BENCH --no-adjoint > $OUT/prof_$K.log 2>&1
done
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:force_adj_gather_kernel -s 10 -c 1 -o $OUT/prof_force_adj_gather_kernel -f $BENCH \
    > $OUT/prof_force_adj_gather_kernel.log 2>&1
ls -la $OUT
