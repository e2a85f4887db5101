#!/bin/bash
# round 2, call 3D: segment length / side inputs per segment vs the period of the main-stem pipeline
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
for sc in 10 6 4 3; do
for sp in "64,0,16,8,8" "64,12,16,8,8" "64,8,16,8,8" "64,6,16,8,8" "64,4,16,8,8"; do
  echo -n "side_cap=$sc sched=$sp: "; TXH_SIDE_CAP=$sc timeout 300 python tools/time_route.py --reps 8 --sched $sp 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_best'], d['ms_mean'], d['sched']['n_tasks'], d['sched']['n_spine'])"
done
done
