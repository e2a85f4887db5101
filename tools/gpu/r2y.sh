#!/bin/bash
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
rm -f gpurun_out/r2y_ss_trace.txt
TXH_SS_TRACE=gpurun_out/r2y_ss_trace.txt timeout 600 python tools/time_enkf.py 2>&1 | tail -1
tail -18 gpurun_out/r2y_ss_trace.txt | head -3
