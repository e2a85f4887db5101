#!/bin/bash
# round 2, call 3H: lane kernel with two rows per thread (deterministic runs)
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=8000
echo "== configs"; timeout 600 python tests/perf/run_configs.py c1 c2 c4 2>&1 | tail -3 | cut -c1-330
echo "== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-250 | head -30
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
