#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
O=gpurun_out
for B in 16 8; do
  for NS in 0 100000; do
    SASSD_TMA_NSPLIT_TILES=$NS timeout 600 python bench.py --steps 20 --warmup 3 --batch $B --no-cpu-baseline > $O/ns_b${B}_${NS}.json 2> $O/ns_b${B}_${NS}.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/ns_b${B}_${NS}.json").read().strip().splitlines()[-1])
    st=d["stages_ms"]
    print("B=$B nsplit_tiles=$NS value %.1f ms %.4f e2e %.1f dense3x3 %.4f 1x1 %.4f first %.4f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],st.get("conv2d_tma[taps=9 256->256]",0),st.get("conv2d_tma[taps=1 256->256]",0),st.get("conv2d_tma[taps=9 320->256]",0)))
except Exception as e: print("B=$B NS=$NS ERR",e)
PY
  done
done
for NS in 0 100000; do
for DBG in 0 8 16 24; do
  echo "== NSPLIT $NS DBG $DBG"
  SASSD_TMA_NSPLIT_TILES=$NS SASSD_TMA_DBG=$DBG timeout 300 python tests/tools/tc_check.py tmaperf1 2>&1 | grep -E "tma f16x3|MISMATCH|Error|error" 
done
done
echo "== trace 3x3 B=1 nsplit"
SASSD_TMA_TRACE=4 timeout 300 python tests/tools/tc_check.py tmaperf1 2>&1 | grep -E "TMA_TRACE" | head -8
echo "== trace 3x3 B=1 no nsplit"
SASSD_TMA_NSPLIT_TILES=0 SASSD_TMA_TRACE=4 timeout 300 python tests/tools/tc_check.py tmaperf1 2>&1 | grep -E "TMA_TRACE" | head -8
