#!/bin/bash
# round 2, call H: full suite on the settled lane defaults, the ctas=4 failure in detail, bench N=1 (EnKF tail without gain store)
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
echo "== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-250 | head -40
echo "== ctas 4 failure"; TXH_LANE_CTAS=4 timeout 600 python -m pytest tests/test_gpu_routing.py -q -m gpu -k texas_scale_short --tb=short 2>&1 | grep -E "^E  |Error|passed|failed" | cut -c1-300 | head -20
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "== bench n=1"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -c 400 gpurun_out/r2h_bench.err; cut -c1-400 gpurun_out/r2h_bench.json
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/r2h_ncu.log 2>&1; tail -2 gpurun_out/r2h_ncu.log | cut -c1-300
