#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 )); }
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log ); echo "== [$(el)s] $name: $(tail -1 gpurun_out/$name.log)"; }
run phases 400 python tests/experimental/time_mlp_update_phases.py
run rollout 200 python -m pytest tests/test_gpu_rollout.py -x -q -s
run t_optim 200 python -m pytest tests/test_gpu_optim.py -q
cat gpurun_out/phases.log
grep -n "diag\|passed\|failed" gpurun_out/rollout.log | head -20
tail -5 gpurun_out/t_optim.log
