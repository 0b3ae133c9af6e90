#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log ); echo "== $name: $(tail -1 gpurun_out/$name.log)"; }
run ncu_ro 300 ncu --set full --import-source on --clock-control none -k regex:k_breakout_rollout --launch-skip 2 -c 1 -f -o gpurun_out/prof_k_breakout_rollout_r02 python bench.py --steps 2 --warmup 1 --no-extra-configs --no-cpu-baseline --no-e2e --no-graph
tail -2 gpurun_out/ncu_ro.log | cut -c1-150
run ncu_xt 300 ncu --set full --import-source on --clock-control none -k regex:k_mlp_update_xt --launch-skip 12 -c 1 -f -o gpurun_out/prof_k_mlp_update_xt_r02 python tests/experimental/check_mlp_update_fused.py --variant 2
tail -2 gpurun_out/ncu_xt.log | cut -c1-150
