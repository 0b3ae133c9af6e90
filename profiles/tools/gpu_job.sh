#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 )); }
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log ); echo "== [$(el)s] $name: $(tail -1 gpurun_out/$name.log)"; }
run probe 100 python tests/experimental/check_umma_probe.py
run rollout 200 python -m pytest tests/test_gpu_rollout.py -x -q -s
grep -n "layout1\|core matri\|mixed\|no-swizzle\|layout type 1" gpurun_out/probe.log
grep -n "diag\|passed\|failed" gpurun_out/rollout.log | head -20
