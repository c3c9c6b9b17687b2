#!/bin/bash
# Round-2 GPU session 21 (1 GPU): the reference's default invocation (18 real frames) on both arms on the same box: CPU arm over all 20
# rounds, then the GPU line with its same-work check against it.
set -x
O=gpurun_out/s21; mkdir -p $O
timeout 300 python bench.py --impl reference --config real --steps 20 --warmup 0 > $O/bench_real_reference_arm.json 2> $O/bench_real_reference_arm.err
timeout 300 python bench.py --config real > $O/bench_real.json 2> $O/bench_real.err
ls -la $O
