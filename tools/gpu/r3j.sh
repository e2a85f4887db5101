#!/bin/bash
# round 2, call 3J: lane kernel, producer lag and rows per region on the final build
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=8000
for lag in 24 32 48; do
  echo -n "lag=$lag: "; TXH_LANE_LAG=$lag timeout 300 python tests/perf/run_configs.py c2 c4 2>&1 | tail -2 | python -c "
import json,sys
print([ (json.loads(l)['gpu_ms']) for l in sys.stdin])"
done
for cap in 512 640 768; do
  echo -n "cap=$cap: "; TXH_LANE_CAP=$cap timeout 300 python tests/perf/run_configs.py c2 c4 2>&1 | tail -2 | python -c "
import json,sys
print([ (json.loads(l)['gpu_ms']) for l in sys.stdin])"
done
