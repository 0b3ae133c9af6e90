#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log ); echo "== $name: $(tail -1 gpurun_out/$name.log)"; }
for c in cmk2048 cmk128 cmm128 cmm2048 cmx128 cmx2048; do
  run probe_$c 60 python tests/experimental/check_umma_probe.py $c
  grep -n "no-swizzle" gpurun_out/probe_$c.log
done
run rollout 200 python -m pytest tests/test_gpu_rollout.py -x -q -s
grep -n "diag\|passed\|failed" gpurun_out/rollout.log | head -20
