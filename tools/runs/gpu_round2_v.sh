#!/bin/bash
# final build, one GPU: the bench exactly as the driver runs it, then smoke()
set -x
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/r2v_bench_n1.json 2> gpurun_out/r2v_bench_n1.err; tail -c 200 gpurun_out/r2v_bench_n1.err
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2v_smoke.txt
