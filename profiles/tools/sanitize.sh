#!/bin/bash
# compute-sanitizer over small instances of the hand-written kernels (run on the B200 box): logs -> gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
for TOOL in memcheck racecheck synccheck; do
  for T in gae loop fused envs; do
    timeout 420 compute-sanitizer --tool $TOOL --print-limit 20 python tests/experimental/sanitize_targets.py $T \
        > gpurun_out/sanitizer_${TOOL}_$T.log 2>&1
    echo "== $TOOL $T rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${TOOL}_$T.log | tail -1)"
  done
done
