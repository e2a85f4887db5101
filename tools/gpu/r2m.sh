#!/bin/bash
# round 2, call M: regular-case bracket shift in the lane kernel, device-side recording in the batched KF path
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
echo "== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-250 | head -30
echo "== configs"; timeout 600 python tests/perf/run_configs.py c1 c2 c4 2>&1 | tail -3 | tee gpurun_out/r2m_configs.jsonl
echo "== batched kf timing"; timeout 600 python tests/perf/time_kf_collection.py 2>&1 | tail -3 | tee gpurun_out/r2m_kf_collection.jsonl
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
