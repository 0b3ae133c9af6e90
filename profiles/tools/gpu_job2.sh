#!/bin/bash
# 2-GPU job: peer all-reduce + one-graph multi-GPU update check, then the bench at N = 1 and N = 2 on the same box
mkdir -p gpurun_out
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log ); echo "== $name: $(tail -1 gpurun_out/$name.log)"; }
run t_optim 300 python -m pytest tests/test_gpu_optim.py -x -q
tail -3 gpurun_out/t_optim.log
run peer2 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu/check_peer_update.py
tail -8 gpurun_out/peer2.log | cut -c1-200
run bench1 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-extra-configs --no-cpu-baseline
grep -o '"value": [0-9.]*, "unit": "agent-steps/s", "n_gpus": [0-9]*, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*' gpurun_out/bench1.log
run bench2 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-extra-configs --no-cpu-baseline
grep -o '"value": [0-9.]*, "unit": "agent-steps/s", "n_gpus": [0-9]*, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*' gpurun_out/bench2.log
