#!/bin/bash
# Round-2 GPU session 17 (2 GPUs): config 3 with the round's final build (scaling table), test_gpu_multi.
set -x
O=gpurun_out/s17; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_c3_2gpu.json 2> $O/bench_c3_2gpu.err
timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > $O/pytest_multi.log 2>&1; echo "rc=$?" >> $O/pytest_multi.log
ls -la $O
