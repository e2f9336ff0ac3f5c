#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_baseline_shapes.py 2>&1 | tail -60 | tee gpurun_out/r2b_pytest.txt
KS_TIMELINE=gpurun_out/r2b_tl timeout 600 python tools/pipe_ab.py 1000000 1:f16 0:f16 2>&1 | tee gpurun_out/r2b_pipe_ab.txt
