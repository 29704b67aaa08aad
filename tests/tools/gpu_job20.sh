#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -n 3 $O/pytest_gpu.log
bash tests/tools/gpu_job19.sh 2>&1 | grep -v "depth [268]"
timeout 600 python bench.py --steps 20 --warmup 3 --batch 16 --no-cpu-baseline > $O/j20_b16.json 2> $O/j20_b16.err
python - <<PY
import json
d=json.loads(open("$O/j20_b16.json").read().strip().splitlines()[-1])
st=d["stages_ms"]
print("B=16 value %.1f ms %.4f e2e %.1f dense3x3 %.4f 1x1 %.4f first %.4f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],st.get("conv2d_tma[taps=9 256->256]",0),st.get("conv2d_tma[taps=1 256->256]",0),st.get("conv2d_tma[taps=9 320->256]",0)))
PY
