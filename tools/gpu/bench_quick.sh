#!/bin/bash
# quick A/B: headline bench (device-resident + e2e), no CPU baseline, no extras
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu,power.draw --format=csv,noheader
timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'window', d['roofline']['kernel_ms_per_launch'], d.get('clocks'))"
