#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 300 python tests/tools/tc_check.py split > $O/tc_split.txt 2>&1; echo "split rc=$?" >> $O/tc_split.txt
timeout 300 python tests/tools/tc_check.py splitperf > $O/tc_splitperf.txt 2>&1
SASSD_SPS_TRACE=2 timeout 300 python tests/tools/tc_check.py splittrace 120000 > $O/tc_trace_big.txt 2>&1
timeout 300 python tests/tools/tc_check.py tma > $O/tc_tma.txt 2>&1; echo "tma rc=$?" >> $O/tc_tma.txt
timeout 300 python tests/tools/tc_check.py tmafull > $O/tc_tmafull.txt 2>&1
SASSD_TMA_TRACE=8 timeout 300 python tests/tools/tc_check.py tmaperf > $O/tc_tmaperf_trace.txt 2>&1
timeout 300 python tests/tools/tc_check.py tmaperf > $O/tc_tmaperf.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_b1.json 2> $O/bench_b1.err
timeout 600 python bench.py --steps 20 --warmup 3 --batch 16 --no-cpu-baseline > $O/bench_b16.json 2> $O/bench_b16.err
tail -n 2 $O/tc_split.txt; grep splitperf $O/tc_splitperf.txt; grep SPS_TRACE $O/tc_trace_big.txt | head -4; tail -n 4 $O/tc_tma.txt; grep "MISMATCH" $O/tc_tmafull.txt | head; grep "TMA_TRACE" $O/tc_tmaperf_trace.txt | head -5; grep "tma f16" $O/tc_tmaperf.txt; tail -n 6 $O/pytest_gpu.log
