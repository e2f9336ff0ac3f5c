#!/bin/bash
# round 2, first hardware run of the new pipeline / parity mode: tests, then A/B timing with a timeline
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_baseline_shapes.py 2>&1 | tail -40 | tee gpurun_out/r2a_pytest.txt
KS_TIMELINE=gpurun_out/r2a_tl timeout 600 python tools/pipe_ab.py 1000000 1:f16 0:f16 1:f16x2 2>&1 | tee gpurun_out/r2a_pipe_ab.txt
timeout 600 python -m pytest tests/test_gpu_baseline_shapes.py -q -m gpu -s 2>&1 | tail -30 | tee gpurun_out/r2a_shapes.txt
