#!/bin/bash
# First GPU run of the code written after round 1's GPU budget was spent (run under gpurun, 1 GPU):
#   * pmwd_powspec_bin (csrc/powspec.cu)                      -> tests/test_gpu_spec.py
#   * nx = 2048 x-pass on two-CTA clusters (csrc/xpass16.cu)  -> test_fused_xpass_2048_cluster_variant
#   * nbody_step_host (stream-overlapped host-array step)     -> test_nbody_step_host_matches_nbody_step
# and the timing of the 2048-point x-pass in the y-slab form of one of 8 GPUs, both kernels.
# When green: drop the PMWD_RUN_UNVALIDATED gates, make PMWD_XPASS16_2048 the default.
export PMWD_RUN_UNVALIDATED=1
python -m pytest tests/test_gpu_spec.py tests/test_gpu_gravity.py -x -q -m gpu -k "spec or cluster or step_host" 2>&1 | tail -15
echo "--- radix-4 shared-memory kernel (default today)"
python tools/time_xpass.py 2048 256 2>&1 | tail -1
echo "--- two-CTA cluster register kernel"
PMWD_XPASS16_2048=1 python tools/time_xpass.py 2048 256 2>&1 | tail -1
echo "--- framework baseline on the same GPU (literal torch transcription of the reference step)"
python tools/torch_baseline.py 256 5 2>&1 | tail -1
