#!/bin/bash
# round 2, call 3F (8 GPUs): the driver's scaling run at N = 8 on the final build
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=8000
nvidia-smi -L | wc -l
N=8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r3f_bench_n$N.json 2> gpurun_out/r3f_bench_n$N.err
echo "rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r3f_bench_n$N.err | tail -5 | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3f_bench_n8.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), d['e2e'].get('value'))
m=d.get('members') or {}; c=d.get('c4_basins') or {}
print('members ms', m.get('ms_per_step'), m.get('value'))
print('c4_basins ms', c.get('ms_per_step'), c.get('value'), c.get('load_imbalance'), c.get('reaches_per_rank'))
PY
