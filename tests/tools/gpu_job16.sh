#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -n 3 $O/pytest_gpu.log
for B in 1 2 4; do
  for NS in 296 0 100000; do
    SASSD_TMA_NSPLIT_TILES=$NS timeout 600 python bench.py --steps 20 --warmup 3 --batch $B --no-cpu-baseline > $O/ns_b${B}_${NS}.json 2> $O/ns_b${B}_${NS}.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/ns_b${B}_${NS}.json").read().strip().splitlines()[-1])
    st=d["stages_ms"]
    print("B=$B nsplit_tiles=$NS value %.1f ms %.4f e2e %.1f dense3x3 %.4f nms %.4f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],st.get("conv2d_tma[taps=9 256->256]",0),st.get("sassd_rescore_nms",0)))
except Exception as e: print("B=$B NS=$NS ERR",e)
PY
  done
done
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv2d_tma_kernel<(128|256)>" -s 27 -c 1 -o $O/prof_r2_dense python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity > $O/ncu_full_dense_r2.log 2>&1
tail -n 2 $O/ncu_full_dense_r2.log
