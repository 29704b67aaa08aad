#!/bin/bash
# final check of the round: GPU parity tests, smoke, the bench lines that go to profiles/
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -n 3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2_bench_b1.json 2> $O/r2_bench_b1.err
timeout 600 python bench.py --steps 20 --warmup 3 --batch 16 --no-cpu-baseline > $O/r2_bench_b16.json 2> $O/r2_bench_b16.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_reference_arm.err
for f in $O/r2_bench_b1.json $O/r2_bench_b16.json $O/r2_bench_reference_arm.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.2f ms/step %.3f e2e %.2f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), "launches", d.get("gpu_launches"), "clocks", d.get("clocks"), "roof", d.get("roofline") and (round(d["roofline"]["frac"],3), d["roofline"].get("mma_issue_frac") and round(d["roofline"]["mma_issue_frac"],3), d["roofline"]["traffic"]), "sparse", d.get("roofline_sparse") and round(d["roofline_sparse"]["frac"],3), "cpu", d.get("cpu_baseline") and d["cpu_baseline"].get("value"), "parity", d.get("parity_check") and d["parity_check"]["ok"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
