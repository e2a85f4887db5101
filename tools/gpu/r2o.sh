#!/bin/bash
# round 2, call O: state check after the container was re-created -- gpu suite, smoke, bench (N = 1)
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
echo "== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-250 | head -30
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.err; echo "rc=$?"; tail -3 gpurun_out/r2o_bench_n1.err | cut -c1-300; cut -c1-1500 gpurun_out/r2o_bench_n1.json
