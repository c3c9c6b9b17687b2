#!/bin/bash
# Round-2 GPU session 6 (4 GPUs): BASELINE configs[3] (40 views x 500k, quaternion, mixed) sharded over 4 GPUs; config 3 at 4 GPUs.
set -x
O=gpurun_out/s6; mkdir -p $O
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --config 4 --steps 20 --warmup 3 > $O/bench_c4_4gpu.json 2> $O/bench_c4_4gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 > $O/bench_c3_4gpu.json 2> $O/bench_c3_4gpu.err
ls -la $O
