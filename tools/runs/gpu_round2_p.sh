#!/bin/bash
# one GPU: C1 (three lambdas, full-size oracle), C4 direct and with its featurizer at the full 50K rows, and the ncu launch list of
# the bench command on the final build
set -x
mkdir -p gpurun_out
timeout 600 python tools/other_configs.py c1 c4 c4f --c4-rows 50000 2>&1 | grep '^{' | tee gpurun_out/r2p_other_configs.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-fast-mode --parity-rows 0 > gpurun_out/r2p_bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/r2p_bench_under_ncu.log
