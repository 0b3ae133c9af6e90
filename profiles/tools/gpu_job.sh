#!/bin/bash
# The job the next gpurun call executes (edited between calls; each step under its own timeout, logs into gpurun_out/).
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 )); }
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log ); echo "== [$(el)s] $name: $(tail -1 gpurun_out/$name.log)"; }
opt() { if [ $(el) -lt ${LIMIT:-1150} ]; then run "$@"; else echo "== skipped $1 (out of time)"; fi; }
run probe 100 python tests/experimental/check_umma_probe.py
run fused 240 python tests/experimental/check_mlp_update_fused.py
run rollout 240 python -m pytest tests/test_gpu_rollout.py -x -q
run t_optim 200 python -m pytest tests/test_gpu_optim.py tests/test_gpu_gae.py -q
run bench 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run smoke 120 python -c "import __graft_entry__ as g; g.smoke()"
run t_rest 500 python -m pytest tests/test_gpu_squared.py tests/test_gpu_experience.py tests/test_gpu_configs.py tests/test_gpu_envs.py -q
opt encgemm 100 python tests/experimental/check_enc_gemm_tcgen05.py
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph --no-extra-configs"
opt launches 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv $B
for K in k_mlp_update_fused:4 k_breakout_rollout:2 k_gae_tile:2; do
  NAME=${K%%:*}; SKIP=${K##*:}
  opt ncu_$NAME 150 ncu --set full --clock-control none --import-source on -k regex:$NAME -s $SKIP -c 1 -o gpurun_out/prof_${NAME}_r02 $B
done
opt san_mem 200 compute-sanitizer --tool memcheck --print-limit 20 python tests/experimental/sanitize_targets.py gae fused
for f in probe fused rollout t_optim bench smoke t_rest; do echo "----- $f"; tail -25 gpurun_out/$f.log; done
