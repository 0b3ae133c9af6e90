#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log ); echo "== $name: $(tail -1 gpurun_out/$name.log)"; }
run fused_v2 200 python tests/experimental/check_mlp_update_fused.py --variant 2
grep -n "MISMATCH\|ALL OK\|SOME\|fused update\|rror" gpurun_out/fused_v2.log | head -60
run tiles 100 python tests/experimental/time_mlp_update_tiles.py
tail -10 gpurun_out/tiles.log
