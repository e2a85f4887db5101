#!/bin/bash
# round 2, call Z: polling back-off of the window kernel
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
for nap in "32,256" "0,0" "16,16" "32,32" "32,64" "64,128" "128,512"; do
  echo "nap=$nap"; TXH_WINDOW_NAP=$nap timeout 300 python tools/time_route.py --reps 10 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_best'], d['ms_mean'])"
done
