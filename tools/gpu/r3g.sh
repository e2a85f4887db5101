#!/bin/bash
# round 2, call 3G: ncu --set full of route_lane_kernel on C2 (100k reaches, M = 1, 2,016 steps in one launch)
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=20000
timeout 300 python tools/time_route.py --M 1 --steps 2016 --reps 3 2>&1 | tail -1 | cut -c300-460
timeout 900 ncu --set full --clock-control none --import-source on -k regex:route_lane_kernel -s 2 -c 1 -o gpurun_out/r3g_lane -f python tools/time_route.py --M 1 --steps 2016 --reps 1 > gpurun_out/r3g_ncu.log 2>&1; echo "rc=$?"; ls -la gpurun_out/r3g_lane.ncu-rep
