#!/bin/bash
# exact Gram diagonal in the parity mode: chain-length probe again, then the parity / property tests
set -x
mkdir -p gpurun_out
timeout 400 python tools/split_chunk_probe.py 2048 8192 16384 32768 2>&1 | grep '^{\|rror' | tee gpurun_out/r2t_split_chunk.txt
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -x -q 2>&1 | grep -v "^E  \|^$" | tail -12 | tee gpurun_out/r2t_pytest.txt
