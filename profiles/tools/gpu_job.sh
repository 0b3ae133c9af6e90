#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log ); echo "== $name: $(tail -1 gpurun_out/$name.log)"; }
run t_host 400 python -m pytest tests/test_gpu_experience.py -x -q
tail -15 gpurun_out/t_host.log | cut -c1-200
