#!/bin/bash
# Round-2 GPU session 2: new size/shape parity tests, real-18 tests, new bench.py (both arms, configs 3 / real / 2),
# ncu --set full source-level capture of the default knn_kernel (round 0 and steady rounds).
set -x
mkdir -p gpurun_out/s2
O=gpurun_out/s2
nproc > $O/host.txt; cat /sys/fs/cgroup/cpu.max >> $O/host.txt 2>&1; free -g >> $O/host.txt; lscpu | head -20 >> $O/host.txt
timeout 900 python -m pytest tests/test_gpu_sizes.py tests/test_gpu_real18.py -q -m gpu -x --durations=10 > $O/pytest_new.log 2>&1
echo "pytest rc=$?" >> $O/pytest_new.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/ref_c3_k3.json 2> $O/ref_c3_k3.err
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 python bench.py --config real --steps 20 --warmup 3 > $O/bench_real.json 2> $O/bench_real.err
timeout 300 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu > $O/bench_c2.json 2> $O/bench_c2.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_kernel -s 3 -c 8 -o $O/prof_knn \
  python bench.py --no-cpu --no-mat --no-normals --steps 8 --warmup 3 > $O/ncu_knn.log 2>&1
ls -la $O
