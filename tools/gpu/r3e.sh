#!/bin/bash
# round 2, call 3E: gpu suite (dense-R in-library test added), smoke, bench
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
echo "== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-250 | head -30
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
bash tools/gpu/bench_quick.sh
