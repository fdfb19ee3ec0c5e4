#!/bin/bash
mkdir -p gpurun_out/r2j
for SPLIT in 1 2 4; do
echo "--- PMWD_P2P_CE_SPLIT=$SPLIT"
PMWD_P2P_CE_SPLIT=$SPLIT timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
  tools/time_slab_force.py 512 PMWD_PIPE=2 PMWD_PIPE=1 2>&1 | grep -v Warning | grep "force" 
done | tee gpurun_out/r2j/slab_split_n2.txt
