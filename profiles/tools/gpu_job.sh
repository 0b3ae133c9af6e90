#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log ); echo "== $name: $(tail -1 gpurun_out/$name.log)"; }
run transpose 60 python tests/experimental/check_umma_transpose.py
cat gpurun_out/transpose.log | head -20
run rollout 300 python -m pytest tests/test_gpu_rollout.py -x -q
tail -5 gpurun_out/rollout.log
