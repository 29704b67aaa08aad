"""Where does detect_stream's time go?  Host time in submit() / unpack(), time blocked in collect()'s event wait, for
several depths.  Usage: python tests/tools/e2e_probe.py [batch] [steps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 400
from sassd_b200 import detectors as DET, ops  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
import sassd_b200 as S  # noqa: E402
from sassd_b200 import checkpoint  # noqa: E402
torch.set_num_threads(8)
cfg = S.Config.fromfile(os.path.join(bench.ROOT, "configs", bench.WORKLOAD["config"]))
model, vg, aset = S.build_from_config(cfg, device=str(dev))
checkpoint.load_state_dict_into(model, checkpoint.make_synthetic_state_dict(0, bench.num_classes()))
pool = 8
frames = bench.make_frames(pool * B)
batches = [frames[i * B:(i + 1) * B] for i in range(pool)]
maxpts = ops.next_pow2(max(max(p.shape[0] for p in fb) for fb in batches))

acc = dict(submit=0.0, wait=0.0, unpack=0.0)
_submit, _collect = DET._GraphedStep.submit, DET._GraphedStep.collect


def submit(self, *a, **k):
    t = time.perf_counter(); r = _submit(self, *a, **k); acc["submit"] += time.perf_counter() - t; return r


def collect(self):
    t = time.perf_counter(); self.done.synchronize(); t1 = time.perf_counter()
    r = self.unpack(); t2 = time.perf_counter()
    acc["wait"] += t1 - t; acc["unpack"] += t2 - t1
    return r


DET._GraphedStep.submit, DET._GraphedStep.collect = submit, collect
for depth in (1, 2, 4, 6, 8):
    for _ in model.detect_stream([batches[i % pool] for i in range(2 * depth)], B, maxpts, depth=depth):
        pass
    torch.cuda.synchronize()
    for k in acc:
        acc[k] = 0.0
    t0 = time.perf_counter()
    for _ in model.detect_stream((batches[i % pool] for i in range(STEPS)), B, maxpts, depth=depth):
        pass
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print("B=%d depth %d: %.1f frames/s, per step: wall %.3f ms = submit %.3f + wait %.3f + unpack %.3f + other %.3f" %
          (B, depth, STEPS * B / wall, wall / STEPS * 1e3, acc["submit"] / STEPS * 1e3, acc["wait"] / STEPS * 1e3,
           acc["unpack"] / STEPS * 1e3, (wall - sum(acc.values())) / STEPS * 1e3), flush=True)
