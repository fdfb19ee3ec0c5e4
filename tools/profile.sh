#!/bin/bash
# ncu evidence for bench.py's kernels (run under gpurun, 1 GPU).  Outputs -> gpurun_out/.
# usage: tools/profile.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
BENCH="python bench.py --no-cpu-baseline --e2e-steps 1 --no-adjoint"
# every launch of our kernels + cuFFT with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    -k regex:'pmwd|fft|Radix|radix' -c 2500 --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH \
    > gpurun_out/launches_${TAG}.bench.log 2>&1
# full captures of the top hand-written kernels, late in the run (-s skips earlier launches)
# (the x-pass kernel is data independent: capture it alone, `ncu --set full ... python tools/time_xpass.py 1024`)
for K in scatter_fast_kernel gather3_kernel kick_drift_kernel; do
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k regex:$K -s 45 -c 1 -o gpurun_out/prof_${K}_${TAG} -f $BENCH --steps 47 \
      > gpurun_out/prof_${K}_${TAG}.log 2>&1
done
ls -la gpurun_out/ | grep ${TAG}
