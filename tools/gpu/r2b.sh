#!/bin/bash
# round 2, call B: lane kernel with producer lag; whole GPU suite; sweeps; ncu of the lane kernel
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
echo "== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25
echo "== configs"; timeout 600 python tests/perf/run_configs.py c1 c2 c4 2>&1 | tail -3 | tee gpurun_out/r2b_configs.jsonl
for sm in 4 8 16 64 100000; do echo "== side_min $sm"; TXH_LANE_SIDE_MIN=$sm timeout 300 python tests/perf/run_configs.py c2 2>&1 | tail -1 | tee -a gpurun_out/r2b_sweep.jsonl; done
for cap in 384 512 1024 2048; do echo "== cap $cap"; TXH_LANE_CAP=$cap timeout 300 python tests/perf/run_configs.py c2 2>&1 | tail -1 | tee -a gpurun_out/r2b_sweep.jsonl; done
echo "== windows"; timeout 900 python tests/perf/run_configs.py c2w 2>&1 | tail -5 | tee gpurun_out/r2b_windows.jsonl
echo "== windows (window kernel)"; TXH_ROUTE_KERNEL=window timeout 900 python tests/perf/run_configs.py c2w 2>&1 | tail -5 | tee -a gpurun_out/r2b_windows.jsonl
echo "== ncu lane"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:route_lane --launch-skip 1 --launch-count 1 -o gpurun_out/r2b_lane_c2 -f python tests/perf/run_configs.py c2 > gpurun_out/r2b_ncu.log 2>&1; tail -3 gpurun_out/r2b_ncu.log
