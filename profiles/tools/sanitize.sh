#!/bin/bash
# compute-sanitizer over small instances of the hand-written kernels (run on the B200 box): logs -> gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
one() { TOOL=$1; T=$2; LIM=$3
  timeout $LIM compute-sanitizer --tool $TOOL --print-limit 20 python tests/experimental/sanitize_targets.py $T > gpurun_out/sanitizer_${TOOL}_$T.log 2>&1
  echo "== $TOOL $T rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${TOOL}_$T.log | tail -1)"; }
one memcheck fused 200
one memcheck loop 150
one memcheck gae 100
one memcheck envs 150
one racecheck fused 250
one racecheck gae 100
one synccheck fused 200
