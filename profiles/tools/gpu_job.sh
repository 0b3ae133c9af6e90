#!/bin/bash
# The job the next gpurun call executes (edited between calls; each step under its own timeout, logs into gpurun_out/).
mkdir -p gpurun_out
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log ); echo "== $name: $(tail -1 gpurun_out/$name.log)"; }
run probe 150 python tests/experimental/check_umma_probe.py
run encgemm 200 python tests/experimental/check_enc_gemm_tcgen05.py
run fused 300 python tests/experimental/check_mlp_update_fused.py
run hlt 200 python tests/experimental/check_heads_loss_tail.py
run rollout 300 python -m pytest tests/test_gpu_rollout.py -x -q
run newtests 900 python -m pytest tests/test_gpu_optim.py tests/test_gpu_configs.py tests/test_gpu_gae.py tests/test_gpu_squared.py tests/test_gpu_experience.py -q
run bench 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run smoke 150 python -c "import __graft_entry__ as g; g.smoke()"
tail -40 gpurun_out/probe.log gpurun_out/rollout.log gpurun_out/encgemm.log gpurun_out/fused.log gpurun_out/hlt.log gpurun_out/newtests.log gpurun_out/smoke.log gpurun_out/bench.log
