#!/bin/bash
# round 2, call E: what an iteration of the lane kernel costs (single region, no streams), vote cadence, step records in smem
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
echo "== lane tests"; timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_routing.py -x -q -m gpu --tb=short 2>&1 | tail -3
for v in 1 64; do for cap in 1024 320; do echo "== C1 cap $cap vote $v"; TXH_LANE_CAP=$cap TXH_LANE_VOTE_EVERY=$v timeout 300 python tests/perf/run_configs.py c1 2>&1 | tail -1 | tee -a gpurun_out/r2e.jsonl; done; done
for v in 1 64; do echo "== C2 vote $v"; TXH_LANE_VOTE_EVERY=$v timeout 300 python tests/perf/run_configs.py c2 2>&1 | tail -1 | tee -a gpurun_out/r2e.jsonl; done
for sm in 16 64; do echo "== C2 side_min $sm"; TXH_LANE_SIDE_MIN=$sm timeout 300 python tests/perf/run_configs.py c2 2>&1 | tail -1 | tee -a gpurun_out/r2e.jsonl; done
echo "== C4"; timeout 600 python tests/perf/run_configs.py c4 2>&1 | tail -1 | tee -a gpurun_out/r2e.jsonl
echo "== ncu C1 single region"; TXH_LANE_CAP=1024 timeout 600 ncu --set full --clock-control none --import-source on -k regex:route_lane --launch-skip 1 --launch-count 1 -o gpurun_out/r2e_lane_c1 -f python tests/perf/run_configs.py c1 > gpurun_out/r2e_ncu.log 2>&1; tail -2 gpurun_out/r2e_ncu.log
