#!/bin/bash
# round 2, call Q: phase timeline of the fused small-system kernel
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
TXH_SS_TRACE=gpurun_out/r2q_ss_trace.txt timeout 600 python tools/time_enkf.py 2>&1 | tail -2
tail -20 gpurun_out/r2q_ss_trace.txt
timeout 600 python tools/time_enkf.py 2>&1 | tail -2
