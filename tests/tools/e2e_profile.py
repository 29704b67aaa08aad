"""GPU dev tool: where does the e2e (host-in, host-out) time of a B=1 step go?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import sassd_b200 as S
from sassd_b200 import checkpoint, ops
from sassd_b200.synth import synth_cloud

cfg = S.Config.fromfile(os.path.join(ROOT, "configs", "car_cfg.py"))
model, vg, aset = S.build_from_config(cfg, device="cuda:0")
checkpoint.load_state_dict_into(model, checkpoint.make_synthetic_state_dict(0, 1))
model.set_precision(ops.PREC_TF32X3)
frames = [[synth_cloud(i)] for i in range(8)]
g = model.enable_cuda_graph(1, 32768)
for i in range(5):
    model.forward_points(frames[i % 8])
torch.cuda.synchronize()
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
N = 40
t_all = time.perf_counter()
for i in range(N):
    f = frames[i % 8]
    t0 = time.perf_counter(); hp, ho, counts = model.stage_points(f); tick("stage", t0)
    total = sum(counts)
    t0 = time.perf_counter(); g.points[:total].copy_(hp[:total], non_blocking=True); g.pt_off.copy_(ho, non_blocking=True); tick("h2d_issue", t0)
    t0 = time.perf_counter(); g.graph.replay(); tick("replay_issue", t0)
    t0 = time.perf_counter(); g.h_det.copy_(g.det, non_blocking=True); g.h_nd.copy_(g.d_ndet, non_blocking=True); g.h_status.copy_(g.status, non_blocking=True); tick("d2h_issue", t0)
    t0 = time.perf_counter(); torch.cuda.current_stream().synchronize(); tick("sync", t0)
    t0 = time.perf_counter(); det, n = g.h_det.numpy(), g.h_nd.numpy(); k = int(n[0]); b = det[0, :k, :7].copy(); tick("unpack", t0)
t_all = time.perf_counter() - t_all
print("per frame ms (usable cores see bench):", {k: round(1e3 * v / N, 3) for k, v in T.items()}, "total %.3f ms" % (1e3 * t_all / N))
# back-to-back replays without host sync in between
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(N):
    g.graph.replay()
e1.record(); torch.cuda.synchronize()
print("back-to-back replay: %.3f ms/step" % (e0.elapsed_time(e1) / N))
t0 = time.perf_counter()
for i in range(N):
    out = model.forward_points(frames[i % 8])
print("forward_points: %.3f ms/frame" % (1e3 * (time.perf_counter() - t0) / N))
t0 = time.perf_counter()
n = 0
for out in model.detect_stream((frames[i % 8] for i in range(N)), 1, 32768):
    n += 1
print("detect_stream: %.3f ms/frame" % (1e3 * (time.perf_counter() - t0) / N))
sys.path.insert(0, ROOT)
import bench
print("usable cores", bench.usable_cores(), "cpu_count", os.cpu_count(), "torch threads", torch.get_num_threads(), "pinned", model._pinned[0].is_pinned())
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,pstate", "--format=csv"], capture_output=True, text=True).stdout)
