#!/bin/bash
# Round-2 GPU session 15 (1 GPU): certificates written against a slightly larger prune bound (margins no longer limited by boxes that
# sit just outside the exact bound): parity + config 3 / real / config 2 lines + launch list.
set -x
O=gpurun_out/s15; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_corr.py -q -m gpu -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python bench.py --no-cpu --no-mat --no-normals --steps 20 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 python bench.py --config real --steps 20 --warmup 3 --no-cpu --no-mat > $O/bench_real.json 2> $O/bench_real.err
timeout 300 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu --no-mat > $O/bench_c2.json 2> $O/bench_c2.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 -c 3000 --csv --log-file $O/launches.csv python bench.py --no-cpu --no-mat --no-normals --no-replay --steps 20 --warmup 3 > $O/ncu_launches.log 2>&1
ls -la $O
