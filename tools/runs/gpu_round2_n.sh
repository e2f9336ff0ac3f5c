#!/bin/bash
# after the staged-upload change: whole GPU suite, the bench as the driver runs it, C1 with its full-size oracle, C4 at its full 50K rows
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | grep -v "^E  \|^$" | tail -15 | tee gpurun_out/r2n_pytest.txt
timeout 900 python bench.py > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err; tail -c 300 gpurun_out/r2n_bench_n1.err
timeout 600 python tools/other_configs.py c1 c4 c4f --c4-rows 50000 2>&1 | grep '^{\|Error\|error' | tee gpurun_out/r2n_other_configs.txt
