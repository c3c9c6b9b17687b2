#!/bin/bash
# Round-2 GPU session 4: full GPU suite after the NN clean-up, config 3 with the far-round rule, configs 4 and 5 on one GPU (breakdown).
set -x
O=gpurun_out/s4; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log
timeout 300 python bench.py --no-cpu --steps 20 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 python bench.py --no-cpu --steps 20 --warmup 3 --flags 16 > $O/bench_c3_noobb.json 2> $O/bench_c3_noobb.err
timeout 900 python bench.py --config 4 --no-cpu --no-mat --no-normals --steps 10 --warmup 3 > $O/bench_c4_1gpu.json 2> $O/bench_c4_1gpu.err
timeout 1200 python bench.py --config 5 --no-cpu --no-mat --no-normals --steps 10 --warmup 3 > $O/bench_c5_1gpu.json 2> $O/bench_c5_1gpu.err
ls -la $O
