#!/bin/bash
# ncu evidence of the round's final kernels + the 3-class / density-sweep lines at one GPU
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv2d_tma_kernel -s 44 -c 9 -o $O/prof_r2_dense -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity > $O/ncu_full_dense_r2.log 2>&1
tail -n 2 $O/ncu_full_dense_r2.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 300 --csv --log-file $O/launches_r2_b1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity > $O/ncu_launch_r2.log 2>&1
for D in mix 5k 120k; do
  timeout 600 python bench.py --config multi_cfg.py --density $D --steps 20 --warmup 3 --no-cpu-baseline > $O/r2_multi_cfg_${D}_n1.json 2> $O/r2_multi_cfg_${D}_n1.err
done
for D in 5k 40k 120k; do
  timeout 600 python bench.py --density $D --steps 20 --warmup 3 --no-cpu-baseline > $O/r2_car_cfg_${D}_n1.json 2> $O/r2_car_cfg_${D}_n1.err
done
for f in $O/r2_multi_cfg_*_n1.json $O/r2_car_cfg_*_n1.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.1f ms/step %.3f e2e %.1f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), d["config"]["workload"][:80], d.get("parity_check") and d["parity_check"]["ok"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
