#!/bin/bash
# round 2, call 3A: batched side-input prefetch of the segments
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
timeout 300 python tools/time_route.py --reps 10 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('window M=64 12 steps', d['ms_best'], d['ms_mean'])"
echo "== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-250 | head -30
bash tools/gpu/bench_quick.sh
