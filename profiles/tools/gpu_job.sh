#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log ); echo "== $name: $(tail -1 gpurun_out/$name.log)"; }
run rollout 300 python -m pytest tests/test_gpu_rollout.py -x -q -s
grep -n "diag\|passed\|failed\|Error" gpurun_out/rollout.log | head -20
run bench 400 python bench.py --steps 10 --warmup 3 --no-extra-configs --no-cpu-baseline --no-e2e
python - <<'PY'
import json
for line in open('gpurun_out/bench.log'):
    if line.startswith('{"metric"'):
        d=json.loads(line); rk=d['roofline_kernels']; print(d['value'], d['ms_per_step'], {k:rk[k]['avg_launch_us'] for k in ('rollout','mlp_update')})
PY
