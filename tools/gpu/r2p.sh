#!/bin/bash
# round 2, call P: fused small-system cluster kernel + packed inflow-rebuild records
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
echo "== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-250 | head -30
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== bench fused"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err; echo "rc=$?"; tail -3 gpurun_out/r2p_bench_n1.err | cut -c1-300; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity_max_rel_err','gpu_launches') if k in d}, d['e2e'], d['roofline']['kernel_ms_per_launch'])
PY
echo "== bench unfused"; TXH_ENKF_FUSED=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d.get('parity_max_rel_err'))"
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 1 --warmup 1 --days 0.25 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/r2p_ncu.log 2>&1; echo "rc=$?"
python tools/launch_summary.py gpurun_out/r2p_launches.csv 2>&1 | tail -25
