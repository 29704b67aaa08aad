#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -n 4 $O/pytest_gpu.log
for B in 1 16; do
timeout 600 python bench.py --steps 20 --warmup 3 --batch $B --no-cpu-baseline > $O/j27_b$B.json 2> $O/j27_b$B.err
python - <<PY
import json
d=json.loads(open("$O/j27_b$B.json").read().strip().splitlines()[-1])
st=d["stages_ms"]
print("B=$B value %.1f ms %.4f e2e %.1f dense3x3 %.4f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],st.get("conv2d_tma[taps=9 256->256]",0)))
print("   sparse frac %.3f ms %.4f"%(d["roofline_sparse"]["frac"], d["roofline_sparse"]["ms_per_step"]), {k:v for k,v in st.items() if not k.startswith("spconv") and not k.startswith("conv2d")}, d["clocks"], d["parity_check"] and d["parity_check"]["ok"], d["roofline"]["traffic"])
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 300 --csv --log-file $O/launches_r2_b1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity > $O/ncu_launch_r2.log 2>&1
python tests/tools/launch_list_md.py $O/launches_r2_b1.csv 3 | head -40
