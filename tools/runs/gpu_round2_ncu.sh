#!/bin/bash
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:gemm_kmajor_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/r2_proj python tools/one_fit.py f16 > gpurun_out/r2_ncu_proj.log 2>&1
timeout 600 $NCU -k regex:gram2_tn_kernel --launch-skip 0 --launch-count 2 -f -o gpurun_out/r2_gram python tools/one_fit.py f16 > gpurun_out/r2_ncu_gram.log 2>&1
timeout 600 $NCU -k regex:gemm2_kmajor_kernel --launch-skip 0 --launch-count 1 -f -o gpurun_out/r2_update python tools/one_fit.py f16 > gpurun_out/r2_ncu_update.log 2>&1
timeout 600 $NCU -k regex:chol_solve_kernel --launch-skip 0 --launch-count 1 -f -o gpurun_out/r2_solve python tools/one_fit.py f16 > gpurun_out/r2_ncu_solve.log 2>&1
ls -la gpurun_out/*.ncu-rep
