#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2_bench_b1.json 2> $O/r2_bench_b1.err
timeout 900 python bench.py --steps 20 --warmup 3 --batch 16 --no-cpu-baseline > $O/r2_bench_b16.json 2> $O/r2_bench_b16.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_reference_arm.err
timeout 600 python bench.py --impl reference-gpu --steps 40 --warmup 3 > $O/r2_bench_reference_gpu_b1.json 2> $O/r2_bench_reference_gpu_b1.err
timeout 600 python bench.py --impl reference-gpu --steps 10 --warmup 2 --batch 16 > $O/r2_bench_reference_gpu_b16.json 2> $O/r2_bench_reference_gpu_b16.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 300 --csv --log-file $O/launches_r2_b1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity > $O/ncu_launch_r2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv2d_tma_kernel -s 12 -c 1 -o $O/prof_r2_dense python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity > $O/ncu_full_dense_r2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spconv_split_kernel -s 36 -c 1 -o $O/prof_r2_sparse_b16 python bench.py --steps 3 --warmup 3 --batch 16 --no-cpu-baseline --no-graph --no-parity > $O/ncu_full_sparse_r2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spconv_split_kernel -s 36 -c 1 -o $O/prof_r2_sparse_b1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity > $O/ncu_full_sparse_b1_r2.log 2>&1
tail -n 3 $O/pytest_gpu.log; cat $O/r2_bench_b1.json | cut -c1-600
