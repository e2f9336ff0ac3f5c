#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "convolver or cifar or fft" 2>&1 | grep -v "^E  \|^$" | tail -40 | tee gpurun_out/r2g_pytest.txt
timeout 300 python tools/solve_probe.py 2>&1 | tee gpurun_out/r2g_solve_probe.txt
KS_SOLVE_NC=16 timeout 300 python tools/solve_probe.py 2>&1 | tee gpurun_out/r2g_solve_probe_nc16.txt
