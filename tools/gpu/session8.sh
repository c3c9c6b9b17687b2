#!/bin/bash
# Round-2 GPU session 8 (1 GPU): neighbour-list NN path -- parity, A/B (flags 0 vs 8), ncu; real-18 and config 5 on one GPU.
set -x
O=gpurun_out/s8; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_corr.py tests/test_gpu_lm.py tests/test_gpu_sizes.py tests/test_gpu_real18.py tests/test_gpu_normals.py -q -m gpu -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python bench.py --no-cpu --steps 20 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 python bench.py --no-cpu --no-mat --no-normals --steps 20 --warmup 3 --flags 8 > $O/bench_c3_noadj.json 2> $O/bench_c3_noadj.err
timeout 300 python bench.py --config real --steps 20 --warmup 3 --no-cpu > $O/bench_real.json 2> $O/bench_real.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --no-cpu --no-mat --no-normals --steps 8 --warmup 3 > $O/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:knn_kernel -s 9 -c 2 -o $O/prof_knn python bench.py --no-cpu --no-mat --no-normals --steps 8 --warmup 3 > $O/ncu_knn.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lm_step_kernel -s 12 -c 3 -o $O/prof_step python bench.py --no-cpu --no-mat --no-normals --steps 8 --warmup 3 > $O/ncu_step.log 2>&1
timeout 1200 python bench.py --config 5 --no-cpu --no-mat --no-normals --steps 10 --warmup 3 > $O/bench_c5_1gpu.json 2> $O/bench_c5_1gpu.err
ls -la $O
