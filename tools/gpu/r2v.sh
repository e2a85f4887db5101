#!/bin/bash
# round 2, call V: fused load with 14-row tasks (15 warps + T) vs 16-row tasks (14 warps + T)
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
for pc in 16 14 13 12; do
echo "== bench TXH_POCKET_CAP=$pc"; TXH_POCKET_CAP=$pc timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['kernel_ms_per_launch'])"
done
