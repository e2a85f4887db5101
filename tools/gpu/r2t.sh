#!/bin/bash
# round 2, call T: window kernel vs warps per CTA and rows per task (room for a shared-memory copy of T?)
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
for w in 16 15 14 13; do
  echo "warps=$w"; TXH_WINDOW_WARPS=$w timeout 300 python tools/time_route.py --reps 10 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_best'], d['ms_mean'], d['sched']['n_tasks'])"
done
for sc in "64,0,14,8,8" "64,0,12,8,8" "64,14,14,8,8" "64,12,12,8,8"; do
  for w in 16; do
  echo "sched=$sc warps=$w"; TXH_WINDOW_WARPS=$w timeout 300 python tools/time_route.py --reps 10 --sched $sc 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_best'], d['ms_mean'], d['sched']['n_tasks'])"
  done
done
