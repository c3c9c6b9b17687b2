#!/bin/bash
# Round-2 GPU session 7 (8 GPUs): BASELINE configs[4] (64 views x 1M) sharded over 8 GPUs; config 3 at 8 GPUs.
set -x
O=gpurun_out/s7; mkdir -p $O
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 20 --warmup 3 > $O/bench_c3_8gpu.json 2> $O/bench_c3_8gpu.err
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --config 5 --steps 20 --warmup 3 > $O/bench_c5_8gpu.json 2> $O/bench_c5_8gpu.err
ls -la $O
