#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
O=gpurun_out
SASSD_PDL=1 timeout 300 python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline --no-parity > $O/j28_pdl.json 2> $O/j28_pdl.err
python - <<PY
import json
d=json.loads(open("$O/j28_pdl.json").read().strip().splitlines()[-1])
print("PDL=1: B=1 value %.1f e2e %.1f"%(d["value"],d["e2e"]["value"]))
PY
for V in "" "SASSD_TMA_DBG=3" "SASSD_TMA_PAIR=1"; do
  echo "== tmaperf $V"
  env $V timeout 200 python tests/tools/tc_check.py tmaperf 2>&1 | grep -E "tma f16x3|MISMATCH"
done
