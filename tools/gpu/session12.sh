#!/bin/bash
# Round-2 GPU session 12 (1 GPU): guessed median select (window around the previous median checked by the NN kernel's epilogue):
# parity, A/B against the three-pass select (flag 64), launch list of one run, DRAM traffic of the steady NN kernel.
set -x
O=gpurun_out/s12; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_corr.py tests/test_gpu_lm.py -q -m gpu -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python bench.py --no-cpu --no-mat --no-normals --steps 20 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 python bench.py --no-cpu --no-mat --no-normals --steps 20 --warmup 3 --flags 64 > $O/bench_c3_noguess.json 2> $O/bench_c3_noguess.err
timeout 300 python bench.py --config real --steps 20 --warmup 3 --no-cpu --no-mat > $O/bench_real.json 2> $O/bench_real.err
timeout 300 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu --no-mat > $O/bench_c2.json 2> $O/bench_c2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv python bench.py --no-cpu --no-mat --no-normals --no-replay --steps 20 --warmup 3 > $O/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'^knn_kernel' -s 12 -c 1 -o $O/prof_knn_steady python bench.py --no-cpu --no-mat --no-normals --steps 20 --warmup 3 > $O/ncu_knn.log 2>&1
ncu -i $O/prof_knn_steady.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct > $O/knn_steady_raw.csv 2>&1
ls -la $O
