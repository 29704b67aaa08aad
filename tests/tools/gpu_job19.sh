#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
O=gpurun_out
echo "== trace whole tiles"
SASSD_TMA_TRACE=4 timeout 300 python tests/tools/tc_check.py tmaperf1 2>&1 | grep -E "TMA_TRACE|tma f16x3" | head -12
echo "== trace half units"
SASSD_TMA_NSPLIT_TILES=100000 SASSD_TMA_TRACE=4 timeout 300 python tests/tools/tc_check.py tmaperf1 2>&1 | grep -E "TMA_TRACE|tma f16x3" | head -12
echo "== in-model trace (B=1)"
for N in 20 21; do SASSD_TMA_TRACE=$N timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity 2>&1 | grep TMA_TRACE | head -9; done
timeout 600 python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline > $O/j19_b1.json 2> $O/j19_b1.err
python - <<PY
import json
d=json.loads(open("$O/j19_b1.json").read().strip().splitlines()[-1])
st=d["stages_ms"]
print("B=1 value %.1f ms %.4f e2e %.1f dense3x3 %.4f 1x1 %.4f first %.4f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],st.get("conv2d_tma[taps=9 256->256]",0),st.get("conv2d_tma[taps=1 256->256]",0),st.get("conv2d_tma[taps=9 320->256]",0)))
PY
timeout 600 python tests/tools/e2e_probe.py 1 400 2>&1 | grep -E "depth|Error|error" 
