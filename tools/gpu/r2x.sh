#!/bin/bash
# round 2, call X: Cholesky with lagged forward substitution + inverted diagonal blocks
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
rm -f gpurun_out/r2x_ss_trace.txt
TXH_SS_TRACE=gpurun_out/r2x_ss_trace.txt timeout 600 python tools/time_enkf.py 2>&1 | tail -1
tail -18 gpurun_out/r2x_ss_trace.txt | head -9
echo "== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-250 | head -30
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "rc=$?"; tail -3 gpurun_out/r2x_bench.err | cut -c1-300; python - <<PY
import json
d=json.loads(open('gpurun_out/r2x_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity_max_rel_err','gpu_launches') if k in d}, d['e2e']['ms_per_step'], d['roofline']['kernel_ms_per_launch'])
PY
