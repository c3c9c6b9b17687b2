#!/bin/bash
# Round-2 GPU session 9 (1 GPU): while-while NN loop A/B (flags 0 vs 32), fused general-path passes (real-18), parity subset.
set -x
O=gpurun_out/s9; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_corr.py tests/test_gpu_real18.py tests/test_gpu_lm.py -q -m gpu -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python bench.py --no-cpu --no-mat --no-normals --steps 20 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 python bench.py --no-cpu --no-mat --no-normals --steps 20 --warmup 3 --flags 32 > $O/bench_c3_steploop.json 2> $O/bench_c3_steploop.err
timeout 300 python bench.py --config real --steps 20 --warmup 3 --no-cpu --no-mat > $O/bench_real.json 2> $O/bench_real.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:knn_far -s 1 -c 1 -o $O/prof_far python bench.py --no-cpu --no-mat --no-normals --steps 4 --warmup 3 > $O/ncu_far.log 2>&1
ls -la $O
