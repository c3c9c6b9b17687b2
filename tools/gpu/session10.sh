#!/bin/bash
# Round-2 GPU session 10 (1 GPU): quaternion-drift regression, neighbour lists inside the far kernel, far-rounds A/B (flags bits 8-11),
# lm_step_kernel phase stamps, DRAM traffic of the current knn / lm_eval kernels (ncu --set full).
set -x
O=gpurun_out/s10; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_corr.py tests/test_gpu_lm.py -q -m gpu -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
for n in 20 40 64; do MVICP_STEP_PROFILE=1 timeout 120 python tools/step_profile.py $n 20000 >> $O/step_profile.txt 2>&1; done
timeout 300 python bench.py --no-cpu --no-mat --no-normals --steps 20 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 python bench.py --no-cpu --no-mat --no-normals --steps 20 --warmup 3 --flags $((4<<8)) > $O/bench_c3_far4.json 2> $O/bench_c3_far4.err
timeout 300 python bench.py --no-cpu --no-mat --no-normals --steps 20 --warmup 3 --flags $((6<<8)) > $O/bench_c3_far6.json 2> $O/bench_c3_far6.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'^knn_kernel|lm_eval_kernel' -s 40 -c 4 -o $O/prof_steady python bench.py --no-cpu --no-mat --no-normals --steps 12 --warmup 3 > $O/ncu_steady.log 2>&1
ncu -i $O/prof_steady.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread > $O/steady_raw.csv 2>&1
ls -la $O
