#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_baseline_shapes.py -s -k "not full_size" 2>&1 | grep -v "^E  \|^$" | tail -70 | tee gpurun_out/r2c_pytest.txt
KS_TIMELINE=gpurun_out/r2c_tl timeout 600 python tools/pipe_ab.py 1000000 1:f16 1:f16:custom_solve=0 2:f16 1:f16x2 2>&1 | tee gpurun_out/r2c_pipe_ab.txt
