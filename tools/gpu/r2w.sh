#!/bin/bash
# round 2, call W: programmatic dependent launch between the window kernel and the small-system kernel
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
echo "== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-250 | head -30
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
for mode in 1 0; do
echo "== bench TXH_PDL=$mode"; TXH_PDL=$mode timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2w_bench_pdl$mode.json 2> gpurun_out/r2w_bench.err; echo "rc=$?"; tail -3 gpurun_out/r2w_bench.err | cut -c1-300; python - <<PY
import json
d=json.loads(open('gpurun_out/r2w_bench_pdl$mode.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity_max_rel_err','gpu_launches') if k in d}, d['e2e']['ms_per_step'], d['roofline']['kernel_ms_per_launch'])
PY
done
