#!/bin/bash
# Round-2 GPU session 18 (4 GPUs): config 3 and config 4 with the round's final build (scaling table).
set -x
O=gpurun_out/s18; mkdir -p $O
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 20 --warmup 3 > $O/bench_c3_4gpu.json 2> $O/bench_c3_4gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 --config 4 --steps 20 --warmup 3 > $O/bench_c4_4gpu.json 2> $O/bench_c4_4gpu.err
ls -la $O
