#!/bin/bash
# round 2, call N (8 GPUs): the driver's scaling runs in miniature -- bench.py at N = 8 and N = 4, reference arm
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=8000
nvidia-smi -L | wc -l
for N in 8 4; do
  echo "== bench $N gpus"
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2n_bench_n$N.json 2> gpurun_out/r2n_bench_n$N.err
  echo "rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r2n_bench_n$N.err | tail -5 | cut -c1-300; cut -c1-200 gpurun_out/r2n_bench_n$N.json
done
echo "== reference arm N=8"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29690 bench.py --impl reference --gpus 8 --steps 1 --warmup 1 2>&1 | grep '"impl"' | cut -c1-400
