#!/bin/bash
# Development helper: one gpurun call = probes + unit checks + parity tests + short bench lines.
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 300 python tests/tools/tc_check.py split > $O/tc_split.txt 2>&1; echo "split rc=$?" >> $O/tc_split.txt
timeout 300 python tests/tools/tc_check.py splitperf > $O/tc_splitperf.txt 2>&1
SASSD_SPS_DBG=16 timeout 300 python tests/tools/tc_check.py splitperf > $O/tc_splitperf_nomma.txt 2>&1
SASSD_SPS_DBG=32 timeout 300 python tests/tools/tc_check.py splitperf > $O/tc_splitperf_nogather.txt 2>&1
SASSD_SPS_DBG=48 timeout 300 python tests/tools/tc_check.py splitperf > $O/tc_splitperf_neither.txt 2>&1
SASSD_SPS_TRACE=2 timeout 300 python tests/tools/tc_check.py splittrace 120000 > $O/tc_trace_big.txt 2>&1
SASSD_SPS_TRACE=2 timeout 300 python tests/tools/tc_check.py splittrace 5300 > $O/tc_trace_small.txt 2>&1
SASSD_SPS_TRACE=2 timeout 300 python tests/tools/tc_check.py splittrace 14000 > $O/tc_trace_mid.txt 2>&1
timeout 300 python tests/tools/tc_check.py tma > $O/tc_tma.txt 2>&1; echo "tma rc=$?" >> $O/tc_tma.txt
timeout 300 python tests/tools/tc_check.py tmaperf > $O/tc_tmaperf.txt 2>&1
SASSD_TMA_PAIR=1 timeout 300 python tests/tools/tc_check.py tma > $O/tc_tma_pair.txt 2>&1
SASSD_TMA_PAIR=1 timeout 300 python tests/tools/tc_check.py tmaperf > $O/tc_tmaperf_pair.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > $O/bench_b1.json 2> $O/bench_b1.err
timeout 600 python bench.py --steps 20 --warmup 3 --batch 16 --no-cpu-baseline > $O/bench_b16.json 2> $O/bench_b16.err
tail -n 5 $O/tc_split.txt $O/pytest_gpu.log
