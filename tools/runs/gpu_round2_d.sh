#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "chol_solve or cosine or fft or save_load or materialized" 2>&1 | grep -v "^E  \|^$" | tail -30 | tee gpurun_out/r2d_pytest.txt
KS_TIMELINE=gpurun_out/r2d_tl timeout 600 python tools/pipe_ab.py 1000000 1:f16 1:f16:dyn_tiles=0 1:f16:custom_solve=0 2:f16 2>&1 | tee gpurun_out/r2d_pipe_ab.txt
