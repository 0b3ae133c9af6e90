#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log ); echo "== $name: $(tail -1 gpurun_out/$name.log)"; }
run fused_v2 240 python tests/experimental/check_mlp_update_fused.py --variant 2
grep -n "MISMATCH\|ALL OK\|SOME\|fused update\|rror" gpurun_out/fused_v2.log | head -60
run ncu_xt 400 ncu --set full --import-source on --clock-control none -k regex:k_mlp_update_xt --launch-skip 12 -c 1 -f -o gpurun_out/prof_k_mlp_update_xt_r02 python tests/experimental/check_mlp_update_fused.py --variant 2
tail -3 gpurun_out/ncu_xt.log
