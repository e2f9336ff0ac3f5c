#!/bin/bash
# usage: tools/gpu_retry.sh [--gpus N] TIMEOUT 'command'   -- retries gpurun while the pod answers "transient"/busy (nothing charged)
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GP --timeout "$T" -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  st=$(python -c "import json;print(json.load(open('/root/repo/gpurun_out/.last_call.json')).get('status'))" 2>/dev/null)
  if [ "$st" != "transient" ] && [ "$rc" != "3" ]; then cat /tmp/gpurun_last.log; exit $rc; fi
  sleep 45
done
cat /tmp/gpurun_last.log; exit 3
