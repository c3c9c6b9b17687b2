#!/bin/bash
# Round-2 GPU session 11 (1 GPU): lm_step_kernel phase stamps; single-pass general (non-unit quaternion) LM kernel: parity + real-18 and
# config-4 timings.
set -x
O=gpurun_out/s11; mkdir -p $O
for n in 20 40 64; do MVICP_STEP_PROFILE=1 timeout 120 python tools/step_profile.py $n 20000 >> $O/step_profile.txt 2>&1; done
timeout 600 python -m pytest tests/test_gpu_lm.py tests/test_gpu_real18.py -q -m gpu -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python bench.py --config real --steps 20 --warmup 3 --no-cpu --no-mat > $O/bench_real.json 2> $O/bench_real.err
timeout 600 python bench.py --config 4 --steps 20 --warmup 3 --no-cpu --no-mat > $O/bench_c4_1gpu.json 2> $O/bench_c4_1gpu.err
ls -la $O
