#!/bin/bash
# round 2, call Q: phase timeline of the fused small-system kernel, cluster of 16 vs 8
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
for nc in 16 8; do
rm -f gpurun_out/r2q_ss_trace_$nc.txt
TXH_SS_CLUSTER=$nc TXH_SS_TRACE=gpurun_out/r2q_ss_trace_$nc.txt timeout 600 python tools/time_enkf.py 2>&1 | tail -1
tail -18 gpurun_out/r2q_ss_trace_$nc.txt
TXH_SS_CLUSTER=$nc timeout 600 python tools/time_enkf.py 2>&1 | tail -1
done
echo "== gpu tests (model)"; timeout 1500 python -m pytest tests/test_gpu_model.py -q -m gpu --tb=short -x 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-250 | head -30
