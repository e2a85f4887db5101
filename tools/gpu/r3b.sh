#!/bin/bash
# round 2, call 3B: state of the build after the EnKF fusion work -- gpu suite (new tests), smoke, bench, launch list,
# ncu --set full of the window kernel (update variant) and the small-system kernel, BASELINE configs
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
echo "== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-250 | head -30
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > gpurun_out/r3b_bench_n1.json 2> gpurun_out/r3b_bench_n1.err; echo "rc=$?"; tail -3 gpurun_out/r3b_bench_n1.err | cut -c1-300; cut -c1-400 gpurun_out/r3b_bench_n1.json
echo "== configs"; timeout 600 python tests/perf/run_configs.py c1 c2 c4 c5 2>&1 | tail -4 | tee gpurun_out/r3b_configs.jsonl | cut -c1-300
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3b_launches.csv python bench.py --steps 1 --warmup 1 --days 0.25 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/r3b_ncu.log 2>&1; echo "rc=$?"
python tools/launch_summary.py gpurun_out/r3b_launches.csv 2>&1 | tail -16
echo "== ncu full"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'route_window_kernel|enkf_small_system' -s 3 -c 4 -o gpurun_out/r3b_full -f python bench.py --steps 1 --warmup 1 --days 0.25 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/r3b_ncu_full.log 2>&1; echo "rc=$?"; ls -la gpurun_out/r3b_full.ncu-rep
