#!/bin/bash
# 8-GPU box: C5 at its full size (N = 2M, 132 blocks, 147 classes, rows sharded by class) on 8 ranks, then the bench at N = 4
# (the one rank count of the scaling run not yet exercised by this round's build: 250 solve columns per rank -> clusters of 4 CTAs)
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
  tools/other_configs.py c5 2>&1 | grep '^{\|Error\|error' | tee gpurun_out/r2o_c5_n8.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus 4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_n4.json 2> gpurun_out/r2o_bench_n4.err
tail -c 400 gpurun_out/r2o_bench_n4.err; head -c 600 gpurun_out/r2o_bench_n4.json
