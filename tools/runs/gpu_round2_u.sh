#!/bin/bash
# 2-GPU box, final build: the two-rank parity tests and the bench at N = 2 as the driver launches it (cpu baseline runs at N = 1 only)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | grep -v "^E  \|^$" | tail -8 | tee gpurun_out/r2u_pytest_multi.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
  bench.py --gpus 2 > gpurun_out/r2u_bench_n2.json 2> gpurun_out/r2u_bench_n2.err
tail -c 300 gpurun_out/r2u_bench_n2.err; head -c 300 gpurun_out/r2u_bench_n2.json
