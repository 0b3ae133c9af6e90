#!/bin/bash
# The measurement sequence behind profiles/ (run on the B200 box from the repo root, e.g.
#   gpurun --timeout 1200 -- 'bash profiles/tools/measure.sh r02'
# ).  Writes everything under gpurun_out/ with the given tag; copy what should be judged into profiles/.
# Numbers printed by a run under ncu are never bench values: the bench JSON comes from the un-profiled run.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph"

timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -4 > $OUT/tests_$TAG.txt
cat $OUT/tests_$TAG.txt
timeout 400 python bench.py > $OUT/bench_b200_$TAG.json 2> $OUT/bench_b200_$TAG.err
timeout 300 python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
# every kernel launch of 4 eager steps with its duration (shares, not absolutes: cold cache, serialised)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv \
    --log-file $OUT/launches_$TAG.csv $B > /dev/null 2>&1
# one full capture per hand-written kernel that matters
for K in k_policy_mlp:20 k_mlp_tail_bwd_tma:20 k_breakout:300 k_gae_fast:2 k_ppo_loss:20 k_clip_adam:20; do
    NAME=${K%%:*}; SKIP=${K##*:}
    timeout 150 ncu --set full --clock-control none --import-source on -k regex:$NAME -s $SKIP -c 1 \
        -o $OUT/prof_${NAME}_$TAG $B > /dev/null 2>&1
done
python - <<PY
import json
d = json.load(open("$OUT/bench_b200_$TAG.json"))
print(d["value"], d["ms_per_step"], d["gpu_launches"], d["e2e"])
print(d["roofline"])
print(d["cpu_baseline"])
print(d["clocks"])
PY
ls -la $OUT/*$TAG*
