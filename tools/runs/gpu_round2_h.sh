#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "chol_solve or cifar_loader" 2>&1 | grep -v "^E  \|^$" | tail -30 | tee gpurun_out/r2h_pytest.txt
timeout 300 python tools/solve_probe.py 2>&1 | tee gpurun_out/r2h_solve_probe.txt
