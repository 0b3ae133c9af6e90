#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log ); echo "== $name: $(tail -1 gpurun_out/$name.log)"; }
run t_all 1200 python -m pytest tests -m gpu -x -q
tail -4 gpurun_out/t_all.log
run smoke 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
tail -2 gpurun_out/smoke.log
run bench_full 600 python bench.py
grep -h '"metric"' gpurun_out/bench_full.log | cut -c1-300
run launches 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-extra-configs --no-cpu-baseline --no-e2e
tail -2 gpurun_out/launches.log | cut -c1-120
