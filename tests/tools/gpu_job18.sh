#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -n 3 $O/pytest_gpu.log
for B in 1 4; do
    timeout 600 python bench.py --steps 20 --warmup 3 --batch $B --no-cpu-baseline > $O/j18_b${B}.json 2> $O/j18_b${B}.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/j18_b${B}.json").read().strip().splitlines()[-1])
    st=d["stages_ms"]
    print("B=$B value %.1f ms %.4f e2e %.1f dense3x3 %.4f 1x1 %.4f first %.4f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],st.get("conv2d_tma[taps=9 256->256]",0),st.get("conv2d_tma[taps=1 256->256]",0),st.get("conv2d_tma[taps=9 320->256]",0)), d["roofline"]["kernel"], d["parity_check"])
except Exception as e: print("B=$B ERR",e)
PY
done
echo "== trace whole tiles"
SASSD_TMA_TRACE=4 timeout 300 python tests/tools/tc_check.py tmaperf1 2>&1 | grep -E "TMA_TRACE|tma f16x3" | head -12
echo "== trace half units"
SASSD_TMA_NSPLIT_TILES=100000 SASSD_TMA_TRACE=4 timeout 300 python tests/tools/tc_check.py tmaperf1 2>&1 | grep -E "TMA_TRACE|tma f16x3" | head -12
echo "== in-model trace (B=1, latency graph off: eager) launches 20.."
for N in 20 21; do SASSD_TMA_TRACE=$N timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity 2>&1 | grep TMA_TRACE | head -9; done
