#!/bin/bash
# round 2, call D: lane kernel v3 (register-resident rows); bench N=1 with the new arms
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
echo "== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-250 | head -40
echo "== configs"; timeout 600 python tests/perf/run_configs.py c1 c2 c4 2>&1 | tail -3 | tee gpurun_out/r2d_configs.jsonl
for sm in 8 16 64 100000; do echo "== side_min $sm"; TXH_LANE_SIDE_MIN=$sm timeout 300 python tests/perf/run_configs.py c2 2>&1 | tail -1 | tee -a gpurun_out/r2d_sweep.jsonl; done
for cap in 384 512 896; do echo "== cap $cap"; TXH_LANE_CAP=$cap timeout 300 python tests/perf/run_configs.py c2 2>&1 | tail -1 | tee -a gpurun_out/r2d_sweep.jsonl; done
echo "== windows"; timeout 900 python tests/perf/run_configs.py c2w 2>&1 | tail -5 | tee gpurun_out/r2d_windows.jsonl
echo "== ncu lane"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:route_lane --launch-skip 1 --launch-count 1 -o gpurun_out/r2d_lane_c2 -f python tests/perf/run_configs.py c2 > gpurun_out/r2d_ncu.log 2>&1; tail -2 gpurun_out/r2d_ncu.log
echo "== bench n=1"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 600 gpurun_out/r2d_bench.err; cut -c1-3000 gpurun_out/r2d_bench.json
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2d_bench_ref.json 2> gpurun_out/r2d_bench_ref.err; tail -c 300 gpurun_out/r2d_bench_ref.err; cut -c1-1500 gpurun_out/r2d_bench_ref.json
