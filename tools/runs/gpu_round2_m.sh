#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "cifar or convolver or least_squares_estimator or f16_matches" 2>&1 | tail -4 | tee gpurun_out/r2m_pytest.txt
KS_TIMELINE=gpurun_out/r2m_tl timeout 600 python tools/pipe_ab.py 1000000 1:f16 4:f16 4:f16:custom_solve=1 4:f16x2 1:f16x2 2>&1 | tee gpurun_out/r2m_pipe_ab.txt
timeout 600 python tools/other_configs.py c4f --c4-rows 10000 2>&1 | tail -3 | tee gpurun_out/r2m_c4f.txt
