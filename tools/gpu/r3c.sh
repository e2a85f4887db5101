#!/bin/bash
# round 2, call 3C (2 GPUs): multi-GPU paths after the EnKF fusion work -- the multi-GPU parity test, bench at N = 2,
# reference arm at N = 2
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --tb=short 2>&1 | tail -3 | cut -c1-300
echo "== bench 2 gpus"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r3c_bench_n2.json 2> gpurun_out/r3c_bench_n2.err; echo "rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r3c_bench_n2.err | tail -5 | cut -c1-300; cut -c1-260 gpurun_out/r3c_bench_n2.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3c_bench_n2.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'))
print('members', json.dumps(d.get('members'))[:600])
print('c4_basins', json.dumps(d.get('c4_basins'))[:600])
PY
