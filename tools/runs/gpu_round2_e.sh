#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "cosine or f16 or fft or materialized" 2>&1 | tail -5 | tee gpurun_out/r2e_pytest.txt
KS_TIMELINE=gpurun_out/r2e_tl timeout 900 python tools/pipe_ab.py 1000000 1:f16 1:f16:gram_chunk_rows=8192 1:f16:gram_chunk_rows=4096 3:f16:reserve_sms=32 3:f16:reserve_sms=16 3:f16:reserve_sms=32:gram_chunk_rows=8192 2>&1 | tee gpurun_out/r2e_pipe_ab.txt
