#!/bin/bash
# 2 GPUs: the whole GPU test suite, then the 2-rank bench with cusolver potrs vs the library's DMMA solve kernel
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v "^E  \|^$" | tail -40 | tee gpurun_out/r2i_pytest_all.txt
timeout 120 python tools/solve_probe.py 2>&1 | tee gpurun_out/r2i_solve_probe.txt
for CS in 0 1; do
  KS_CUSTOM_SOLVE=$CS KS_TIMELINE=gpurun_out/r2i_tl_cs$CS timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline --precision f16 --no-e2e > gpurun_out/r2i_bench_n2_cs$CS.json 2> gpurun_out/r2i_bench_n2_cs$CS.err
  tail -c 300 gpurun_out/r2i_bench_n2_cs$CS.err
done
