#!/bin/bash
# Round-2 GPU session 3: first device run of the warp-cooperative NN kernel (coop.cuh): parity + A/B against the per-lane kernel.
set -x
O=gpurun_out/s3; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_corr.py tests/test_gpu_zz_closed_form.py -q -m gpu -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
for F in 0 16 32; do
  timeout 300 python bench.py --no-cpu --no-mat --no-normals --steps 20 --warmup 3 --flags $F > $O/bench_f$F.json 2> $O/bench_f$F.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_coop -s 3 -c 8 -o $O/prof_coop \
  python bench.py --no-cpu --no-mat --no-normals --steps 8 --warmup 3 > $O/ncu_coop.log 2>&1
ls -la $O
