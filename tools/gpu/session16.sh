#!/bin/bash
# Round-2 GPU session 16 (1 GPU): the round's closing validation: whole -m gpu suite, smoke(), the default bench line (with the CPU leg
# and the materialised e2e), the reference arm, DRAM traffic of the certified round's two kernels, config 5 on one GPU at 20 steps.
set -x
O=gpurun_out/s16; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'knn_cert_kernel|knn_todo_kernel' -s 16 -c 2 -o $O/prof_cert python bench.py --no-cpu --no-mat --no-normals --no-replay --steps 20 --warmup 3 > $O/ncu_cert.log 2>&1
ncu -i $O/prof_cert.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active > $O/cert_raw.csv 2>&1
timeout 900 python bench.py --config 5 --steps 20 --warmup 3 --no-cpu --no-mat > $O/bench_c5_1gpu.json 2> $O/bench_c5_1gpu.err
ls -la $O
