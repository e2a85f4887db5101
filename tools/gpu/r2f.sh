#!/bin/bash
# round 2, call F: region timelines of the lane kernel (C2 x3, C4), to explain the run-to-run spread
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
cat > /tmp/one.py <<'PY'
import sys, os, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
from tx_fast_hydrology_b200 import synthetic as S
from tx_fast_hydrology_b200.network import RiverNetwork, Forcing
which, nrep = sys.argv[1], int(sys.argv[2])
if which == "c2": nd, seed, T = S.make_network(100000, 2), 2, 2016
else: nd, seed, T = S.make_network(2_700_000, 3, n_basins=64), 3, 288
n = nd["endnodes"].size
prm = S.make_params(n, seed); t0 = 1_700_000_000 * 10**9
times, table = S.make_forcing(n, T, 300.0, seed, t0_ns=t0)
net = RiverNetwork(nd["endnodes"]); net.compute_coeffs(prm["K"], prm["X"], 300.0)
f = Forcing(net, times, table)
O = net.alloc_state(1); I = net.alloc_state(1)
def reset(): net.pack_host(prm["o_t"][:, None], 1, O); net.init_inflows(O, I, 1)
reset(); net.route_run(O, I, 1, f, t0, int(300e9), T); net.check()
for rep in range(nrep):
    reset(); torch.cuda.synchronize()
    os.environ["TXH_LANE_TRACE"] = f"gpurun_out/lane_{which}_{rep}.bin"
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); net.route_run(O, I, 1, f, t0, int(300e9), T); e1.record(); torch.cuda.synchronize()
    print(which, rep, "ms", e0.elapsed_time(e1), flush=True)
PY
timeout 600 python /tmp/one.py c2 4 2>&1 | tail -5
for r in 0 1 2 3; do python tools/lane_trace.py gpurun_out/lane_c2_$r.bin | head -22; done
timeout 900 python /tmp/one.py c4 2 2>&1 | tail -3
for r in 0 1; do python tools/lane_trace.py gpurun_out/lane_c4_$r.bin | head -40; done
