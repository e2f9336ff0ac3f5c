#!/bin/bash
# 4-GPU box: the bench at N = 4 exactly as the driver launches it (250 solve columns per rank: DMMA solve on clusters of 4 CTAs)
set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus 4 > gpurun_out/r2r_bench_n4.json 2> gpurun_out/r2r_bench_n4.err
tail -c 300 gpurun_out/r2r_bench_n4.err; head -c 400 gpurun_out/r2r_bench_n4.json
