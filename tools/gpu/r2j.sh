#!/bin/bash
# round 2, call J: TMA bulk staging in the window kernel (on/off), whole suite
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
echo "== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-250 | head -30
for b in 1 0 1 0; do echo "== bench bulk=$b"; TXH_WINDOW_BULK=$b timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'window_ms', d['roofline']['kernel_ms_per_launch'])"; done
echo "== configs"; timeout 600 python tests/perf/run_configs.py c1 c2 c4 c5 2>&1 | tail -4 | tee gpurun_out/r2j_configs.jsonl
