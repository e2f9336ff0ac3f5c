#!/bin/bash
# 2 GPUs: multi-rank parity tests, then the bench in both stream arrangements (short: no CPU baseline)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2f_pytest_multi.txt
for P in 1 3; do
  KS_PIPELINE=$P KS_RESERVE_SMS=$([ $P = 3 ] && echo 32 || echo 8) KS_TIMELINE=gpurun_out/r2f_tl_p$P timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2f_bench_n2_p$P.json 2> gpurun_out/r2f_bench_n2_p$P.err
  tail -c 600 gpurun_out/r2f_bench_n2_p$P.err
done
