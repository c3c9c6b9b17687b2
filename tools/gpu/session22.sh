#!/bin/bash
# Round-2 GPU session 22 (1 GPU): bench.py with the timed passes free of instrumentation (value / e2e / materialised e2e: K calls between
# barriers; per-round breakdown from a repeated, instrumented pass).
set -x
O=gpurun_out/s22; mkdir -p $O
timeout 200 python bench.py --no-cpu > $O/bench_default.json 2> $O/bench_default.err
tail -3 $O/bench_default.err
