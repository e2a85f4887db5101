#!/bin/bash
# round 2, call G: CTAs per SM x lag sweep of the lane kernel
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
echo "== lane tests"; timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_routing.py -x -q -m gpu --tb=short 2>&1 | tail -3
for c in 1 2 4 6; do for lag in 24 32 48; do echo "== C2 ctas $c lag $lag"; TXH_LANE_CTAS=$c TXH_LANE_LAG=$lag timeout 300 python tests/perf/run_configs.py c2 2>&1 | tail -1 | tee -a gpurun_out/r2g.jsonl; done; done
for c in 2 4; do for sm in 8 16 64; do echo "== C2 ctas $c side_min $sm"; TXH_LANE_CTAS=$c TXH_LANE_SIDE_MIN=$sm timeout 300 python tests/perf/run_configs.py c2 2>&1 | tail -1 | tee -a gpurun_out/r2g.jsonl; done; done
for c in 1 2 4; do echo "== C4 ctas $c"; TXH_LANE_CTAS=$c timeout 600 python tests/perf/run_configs.py c4 2>&1 | tail -1 | tee -a gpurun_out/r2g.jsonl; done
echo "== C1"; timeout 300 python tests/perf/run_configs.py c1 2>&1 | tail -1 | tee -a gpurun_out/r2g.jsonl
