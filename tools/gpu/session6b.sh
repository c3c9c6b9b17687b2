#!/bin/bash
# Round-2 GPU session 6b (4 GPUs): BASELINE configs[3] (40 views x 500k, quaternion, mixed), 20 rounds, after the quaternion-drift fix.
set -x
O=gpurun_out/s6; mkdir -p $O
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --config 4 --steps 20 --warmup 3 > $O/bench_c4_4gpu.json 2> $O/bench_c4_4gpu.err
tail -5 $O/bench_c4_4gpu.err
ls -la $O
