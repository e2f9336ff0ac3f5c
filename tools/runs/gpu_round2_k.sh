#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -q -m gpu -x -k "two_rank or f16 or cosine or error_paths" 2>&1 | grep -v "^E  \|^$" | tail -12 | tee gpurun_out/r2k_pytest.txt
for LA in 1 2 3; do
  KS_LOOKAHEAD=$LA timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --precision f16 --no-e2e --parity-rows 0 > gpurun_out/r2k_bench_n2_la$LA.json 2> gpurun_out/r2k_bench_n2_la$LA.err
  tail -c 200 gpurun_out/r2k_bench_n2_la$LA.err
done
