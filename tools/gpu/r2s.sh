#!/bin/bash
# round 2, call S: ensemble update applied by the window kernel while it loads its tasks (fused load)
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
echo "== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-250 | head -30
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
for mode in 1 0; do
echo "== bench TXH_ENKF_FUSE_LOAD=$mode"; TXH_ENKF_FUSE_LOAD=$mode timeout 900 python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/r2s_bench_fl$mode.json 2> gpurun_out/r2s_bench.err; echo "rc=$?"; tail -3 gpurun_out/r2s_bench.err | cut -c1-300; python - <<PY
import json
d=json.loads(open('gpurun_out/r2s_bench_fl$mode.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity_max_rel_err','gpu_launches') if k in d}, d['e2e']['ms_per_step'], d['roofline']['kernel_ms_per_launch'])
PY
done
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --steps 1 --warmup 1 --days 0.25 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/r2s_ncu.log 2>&1; echo "rc=$?"
python tools/launch_summary.py gpurun_out/r2s_launches.csv 2>&1 | tail -25
