#!/bin/bash
# Round-2 GPU session 1: first device run of the graph-walk / OBB-far NN paths (bit-identity), A/B of flags 0/8/16/24,
# ncu --set full of knn_walk_kernel and lm_step_kernel.
set -x
mkdir -p gpurun_out/s1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s1/smi.txt
MVICP_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_corr.py -q -m gpu -k "schedule or synthetic" > gpurun_out/s1/pytest_exp.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s1/pytest_exp.log
for F in 0 8 16 24; do
  timeout 300 python bench.py --no-cpu --steps 20 --warmup 3 --flags $F > gpurun_out/s1/bench_f$F.json 2> gpurun_out/s1/bench_f$F.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_walk -s 6 -c 2 -o gpurun_out/s1/prof_walk \
  python bench.py --no-cpu --steps 8 --warmup 3 --flags 8 > gpurun_out/s1/ncu_walk.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lm_step_kernel -s 12 -c 3 -o gpurun_out/s1/prof_step \
  python bench.py --no-cpu --steps 8 --warmup 3 > gpurun_out/s1/ncu_step.log 2>&1
ls -la gpurun_out/s1
