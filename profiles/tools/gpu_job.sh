#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log ); echo "== $name: $(tail -1 gpurun_out/$name.log)"; }
run fused_v2 200 python tests/experimental/check_mlp_update_fused.py --variant 2
grep -n "MISMATCH\|ALL OK\|SOME\|fused update\|rror" gpurun_out/fused_v2.log | head -20
run t_optim 400 python -m pytest tests/test_gpu_optim.py tests/test_gpu_rollout.py -x -q
tail -6 gpurun_out/t_optim.log
run bench 400 python bench.py --steps 10 --warmup 3 --no-extra-configs
grep -o '"value": [0-9.]*, "unit": "agent-steps/s", "n_gpus": [0-9]*, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*' gpurun_out/bench.log
grep -o '"e2e": {[^}]*}' gpurun_out/bench.log; grep -o '"clocks": {[^}]*}' gpurun_out/bench.log
