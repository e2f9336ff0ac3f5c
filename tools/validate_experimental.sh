#!/bin/bash
# First GPU run of the next round: the two code paths written after the round-1 GPU budget was spent
# (split-operand precision mode KS_PRECISION_F16X2, device confusion matrix), then the 4/8-GPU scaling of the column-sharded solve.
#   gpurun --timeout 900 -- 'bash tools/validate_experimental.sh'
set -x
mkdir -p gpurun_out
KS_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "f16x2 or confusion" 2>&1 | tail -20 | tee gpurun_out/pytest_experimental.txt
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/accuracy_f16x2.txt
import sys
sys.argv = ["accuracy_probe.py", "f16", "f16x2"]
sys.path.insert(0, "tools")
import accuracy_probe
accuracy_probe.main(precisions=("f16", "f16x2"))
PY
