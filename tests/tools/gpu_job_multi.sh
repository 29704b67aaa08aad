#!/bin/bash
# bench at N = 1 and N = $1 GPUs (torchrun, NCCL): car_cfg B=1 (the headline config), + multi_cfg density mix at N,
# + (REF=1) the reference arm launched the same way
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/scale_n1.json 2> $O/scale_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $O/scale_n$N.json 2> $O/scale_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --config multi_cfg.py --density mix --steps 20 --warmup 5 --no-cpu-baseline > $O/scale_multi_cfg_mix_n$N.json 2> $O/scale_multi_cfg_mix_n$N.err
if [ "$REF" = "1" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 5 --warmup 1 > $O/scale_ref_n$N.json 2> $O/scale_ref_n$N.err
fi
for f in $O/scale_n1.json $O/scale_n$N.json $O/scale_multi_cfg_mix_n$N.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n", d["n_gpus"], "value %.1f ms/step %.3f e2e %.1f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), d["details"]["timed_region"], d["clocks"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
tail -n 3 $O/scale_n$N.err
