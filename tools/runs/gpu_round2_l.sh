#!/bin/bash
set -x
mkdir -p gpurun_out
for LA in 2 1 3; do
  KS_LOOKAHEAD=$LA timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$LA bench.py --gpus 8 --steps 3 --warmup 2 --no-cpu-baseline --precision f16 --no-e2e --parity-rows 0 > gpurun_out/r2l_bench_n8_la$LA.json 2> gpurun_out/r2l_bench_n8_la$LA.err
  tail -c 200 gpurun_out/r2l_bench_n8_la$LA.err
done
KS_LOOKAHEAD=2 KS_TIMELINE=gpurun_out/r2l_tl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 1 --warmup 1 --no-cpu-baseline --precision f16 --no-e2e --parity-rows 0 --no-fast-mode > gpurun_out/r2l_bench_n8_tl.json 2> gpurun_out/r2l_bench_n8_tl.err
