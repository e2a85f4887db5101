#!/bin/bash
# round 2, call L (2 GPUs): batched KF timing on a balanced split; member sharding with 8 members per GPU; full bench at N=2
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
echo "== batched kf timing"; timeout 600 python tests/perf/time_kf_collection.py 2>&1 | tail -3 | tee gpurun_out/r2l_kf_collection.jsonl
echo "== members, 8 per GPU (Mtot = 16)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 3 --sharding members --members 8 --no-extras --no-e2e > gpurun_out/r2l_members8.json 2> gpurun_out/r2l_members8.err; tail -c 600 gpurun_out/r2l_members8.err; cut -c1-260 gpurun_out/r2l_members8.json
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --tb=short 2>&1 | tail -3 | cut -c1-300
echo "== bench 2 gpus"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2l_bench_n2.json 2> gpurun_out/r2l_bench_n2.err; tail -c 600 gpurun_out/r2l_bench_n2.err; cut -c1-260 gpurun_out/r2l_bench_n2.json
