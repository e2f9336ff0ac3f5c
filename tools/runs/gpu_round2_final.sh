#!/bin/bash
# final 1-GPU record of round 2: the whole GPU test suite, the bench as the driver runs it, the reference arm, the ncu launch list of
# the same bench command, the other BASELINE configurations
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v "^E  \|^$" | tail -15 | tee gpurun_out/r2z_pytest.txt
timeout 900 python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; tail -c 300 gpurun_out/r2z_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --parity-rows 0 > gpurun_out/r2z_bench_under_ncu.log 2>&1
timeout 900 python tools/other_configs.py c2 c4 c4f c5 --c5-rows 400000 --c5-blocks 8 --c4-rows 20000 2>&1 | grep '^{' | tee gpurun_out/r2z_other_configs.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2z_smoke.txt
