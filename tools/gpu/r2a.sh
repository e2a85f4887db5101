#!/bin/bash
# round 2, call A: first contact of the lane kernel with the GPU
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
echo "== lane tests"; timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_routing.py -x -q -m gpu 2>&1 | tail -15
echo "== model tests"; timeout 900 python -m pytest tests/test_gpu_model.py tests/test_collection.py -x -q -m gpu 2>&1 | tail -15
echo "== configs"; timeout 600 python tests/perf/run_configs.py c1 c2 c4 2>&1 | tail -5 | tee gpurun_out/r2a_configs.jsonl
echo "== kf budget"; timeout 300 python tests/perf/kf_error_budget.py 2>&1 | tail -2 | tee gpurun_out/r2a_kf_budget.json
for cap in 320 512 1024 2048; do echo "== cap $cap"; TXH_LANE_CAP=$cap timeout 300 python tests/perf/run_configs.py c2 2>&1 | tail -1 | tee -a gpurun_out/r2a_sweep.jsonl; done
for sm in 4 8 16 64 100000; do echo "== side_min $sm"; TXH_LANE_SIDE_MIN=$sm timeout 300 python tests/perf/run_configs.py c2 2>&1 | tail -1 | tee -a gpurun_out/r2a_sweep.jsonl; done
