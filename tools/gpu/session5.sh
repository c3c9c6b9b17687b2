#!/bin/bash
# Round-2 GPU session 5 (2 GPUs): sharded path -- bit identity tests and the new bench.py under torchrun (1-GPU replay inside).
set -x
O=gpurun_out/s5b; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > $O/pytest_multi.log 2>&1; echo "rc=$?" >> $O/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_c3_2gpu.json 2> $O/bench_c3_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference --single-rounds 0 > $O/ref_c3_2gpu.json 2> $O/ref_c3_2gpu.err
ls -la $O
