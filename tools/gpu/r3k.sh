#!/bin/bash
# round 2, call 3K: gpu suite + smoke + full bench (N = 1) after the coefficient-tracking change
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
echo "== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-250 | head -30
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > gpurun_out/r3k_bench_n1.json 2> gpurun_out/r3k_bench_n1.err; echo "rc=$?"; tail -3 gpurun_out/r3k_bench_n1.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3k_bench_n1.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'c4', d['c4_basins']['ms_per_step'], d['parity_max_rel_err'])
PY
