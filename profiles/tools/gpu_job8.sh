#!/bin/bash
# N-GPU job (N = number of visible GPUs): peer all-reduce check + the bench on all of them
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log ); echo "== $name: $(tail -1 gpurun_out/$name.log)"; }
run peer$N 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu/check_peer_update.py
tail -8 gpurun_out/peer$N.log | cut -c1-200
run bench$N 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-extra-configs --no-cpu-baseline --no-e2e
grep -o '"value": [0-9.]*, "unit": "agent-steps/s", "n_gpus": [0-9]*, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*' gpurun_out/bench$N.log
tail -3 gpurun_out/bench$N.log | cut -c1-300
