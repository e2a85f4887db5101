#!/bin/bash
# round 2, call I (2 GPUs): member-sharded ensemble on the peer-read transform, bench --gpus 2 with the new arms
cd $GRAFT_REPO_ROOT
export TXH_WATCHDOG_MS=4000
nvidia-smi -L | head -4
echo "== routing tests (re-lag fix)"; timeout 900 python -m pytest tests/test_gpu_routing.py tests/test_gpu_configs.py -q -m gpu --tb=short 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-250 | head
echo "== multi-gpu check"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/check_multi_gpu.py 2>&1 | grep -v "^$" | tail -12 | cut -c1-300
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --tb=short 2>&1 | tail -5 | cut -c1-300
echo "== bench 2 gpus"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err; tail -c 1500 gpurun_out/r2i_bench_n2.err; cut -c1-300 gpurun_out/r2i_bench_n2.json
echo "== bench 2 gpus allgather members"; TXH_MEMBER_UPDATE=allgather timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 3 --warmup 3 --sharding members --members 32 --no-extras --no-e2e > gpurun_out/r2i_bench_n2_allgather.json 2> gpurun_out/r2i_bench_n2_allgather.err; tail -c 500 gpurun_out/r2i_bench_n2_allgather.err; cut -c1-300 gpurun_out/r2i_bench_n2_allgather.json
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29537 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | grep '"impl"' | cut -c1-1200
