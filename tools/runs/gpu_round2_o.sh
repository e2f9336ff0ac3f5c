#!/bin/bash
# 8-GPU box: C5 at its full size (N = 2M, 132 blocks, 147 classes, rows sharded by class) on 8 ranks
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
  tools/other_configs.py c5 2>&1 | grep '^{\|Error\|error' | tee gpurun_out/r2o_c5_n8.txt
