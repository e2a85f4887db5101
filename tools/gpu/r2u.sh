#!/bin/bash
# round 2, call U: timeline of a window launch that applies the ensemble update while loading (UPD variant)
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
TXH_TRACE_FILE=gpurun_out/r2u_win_trace.bin timeout 600 python bench.py --steps 1 --warmup 1 --days 0.25 --no-cpu-baseline --no-extras --no-e2e 2>&1 | tail -1 | cut -c1-200
python tools/trace_window.py gpurun_out/r2u_win_trace.bin
python tools/trace_critical.py gpurun_out/r2u_win_trace.bin
