#!/bin/bash
# bench at N = 1 and N = $1 GPUs (torchrun, NCCL), + the reference arm launched the same way
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/scale_n1.json 2> $O/scale_n1.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > $O/scale_n$n.json 2> $O/scale_n$n.err
  fi
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 5 --warmup 1 > $O/scale_ref_n$N.json 2> $O/scale_ref_n$N.err
for f in $O/scale_n*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n", d["n_gpus"], "value %.1f ms/step %.3f e2e %.1f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), d["details"]["timed_region"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
tail -n 3 $O/scale_n$N.err
