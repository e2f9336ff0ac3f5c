#!/bin/bash
# 8 GPUs: the bench as the driver runs it (default settings), then the fast mode alone with cuSOLVER potrs for comparison
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2j_bench_n8.json 2> gpurun_out/r2j_bench_n8.err
tail -c 300 gpurun_out/r2j_bench_n8.err
KS_CUSTOM_SOLVE=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 3 --precision f16 --no-e2e --parity-rows 0 > gpurun_out/r2j_bench_n8_potrs.json 2> gpurun_out/r2j_bench_n8_potrs.err
tail -c 300 gpurun_out/r2j_bench_n8_potrs.err
