#!/bin/bash
# round 2, call K: batched Kalman filters of a collection; AsyncSimulation row-0 quirk; configs with the kernel choice
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
echo "== collection tests"; timeout 900 python -m pytest tests/test_collection.py -q -m gpu --tb=short -s 2>&1 | grep -E "^E  |passed|failed|FAILED|batched=" | cut -c1-300 | head -30
echo "== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-250 | head -30
echo "== configs"; timeout 600 python tests/perf/run_configs.py c1 c2 c4 c5 2>&1 | tail -4 | tee gpurun_out/r2k_configs.jsonl
echo "== batched kf timing"; timeout 600 python tests/perf/time_kf_collection.py 2>&1 | tail -3 | tee gpurun_out/r2k_kf_collection.jsonl
