#!/bin/bash
# 8 GPUs: in-process A/B of the slab force variants (pipelined vs plain transposes)
mkdir -p gpurun_out/r2n
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
  tools/time_slab_force.py 512 PMWD_PIPE=1 PMWD_PIPE=2 PMWD_PIPE=1 PMWD_PIPE=2 PMWD_PIPE=4 2>&1 | grep -v Warning | grep "force" | tee gpurun_out/r2n/slab_ab_n8.txt
PMWD_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 \
  tools/time_slab_force.py 512 PMWD_PIPE=2 2>&1 | grep "start@" | tail -3 | tee gpurun_out/r2n/slab_trace_n8.txt
