#!/bin/bash
# round 2, call R: timeline of a 12-step window launch at the headline size (critical path analysis)
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
TXH_TRACE_FILE=gpurun_out/r2r_win_trace.bin timeout 600 python tools/time_route.py --reps 1 2>&1 | tail -1 | cut -c1-300
python tools/trace_window.py gpurun_out/r2r_win_trace.bin
python tools/trace_critical.py gpurun_out/r2r_win_trace.bin
