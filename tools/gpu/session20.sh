#!/bin/bash
# Round-2 GPU session 20 (1 GPU): page-locked staging of the materialised correspondence lists (mvicp_host_alloc): the C++ drivers and
# the all-edges API through it, and the bench line with the materialised e2e.
set -x
O=gpurun_out/s20; mkdir -p $O
timeout 600 python -m pytest tests/test_app_multiview.py tests/test_gpu_real18.py tests/test_gpu_corr.py -q -m gpu -x -k "driver or real18 or all_edges" > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 400 python bench.py --no-cpu > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python bench.py --config real --no-cpu > $O/bench_real.json 2> $O/bench_real.err
ls -la $O
