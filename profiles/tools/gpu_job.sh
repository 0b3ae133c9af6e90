#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log ); echo "== $name: $(tail -1 gpurun_out/$name.log)"; }
run t_host 400 python -m pytest tests/test_gpu_experience.py -x -q -k "eager_graph_and_host"
tail -6 gpurun_out/t_host.log | cut -c1-200
run bench 400 python bench.py --steps 10 --warmup 3 --no-extra-configs --no-cpu-baseline
python - <<'PY'
import json
for line in open('gpurun_out/bench.log'):
    if line.startswith('{"metric"'):
        d=json.loads(line); print(d['value'], d['ms_per_step'], d['clocks'], d['e2e'])
PY
