#!/usr/bin/env python
"""bench.py — SA-SSD inference hot path on B200: frames/sec on synthetic KITTI-shaped clouds.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the whole hot path (voxelize -> anchors_mask -> 13 sparse convs ->
BEV neck -> heads -> guided anchors -> PSWarp -> rotated NMS) over one batch of B frames
(default B=1 = BASELINE.json configs[1]: car_cfg.py, batch 1, ~20 k points per frame).
Rank 0 prints ONE JSON line (see the driver contract): `value` = whole-job frames/s with
inputs resident in HBM, `e2e` = the same through the public API from host buffers (pinned
H2D of the raw points, D2H of the detections inside the timed region), `roofline` for the
dominant kernel, `roofline_sparse` for the 13 ruled sparse convs (pair-model bytes, SURVEY
§8d), `cpu_baseline` = the CPU oracle port timed on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "KITTI frames/sec (~20k pts, car_cfg voxel grid)"
ORACLE_CFG = dict(voxel_size=[0.05, 0.05, 0.1], pc_range=[0, -40., -3., 70.4, 40., 1.], max_points=5, max_voxels=20000,
                  sparse_shape=[40, 1600, 1408],
                  anchor_cfgs=[dict(sizes=[1.6, 3.9, 1.56], anchor_strides=[0.4, 0.4, 1.0],
                                    anchor_offsets=[0.2, -39.8, -1.78], rotations=[0, 1.57])],
                  grid_offsets=(0., 40.), featmap_stride=.4, score_thr=0.3, iou_thr=0.1)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self._stop, self._t = gpu_index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["unsampled"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                    samples=len(sm))


def usable_cores():
    """Cores this process may really use: affinity mask capped by the cgroup CPU quota (a container on a
    128-core host often has far fewer; oversubscribing torch's thread pools makes every CPU op crawl)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            q, p = f.read().split()
        if q != "max":
            n = min(n, max(1, int(float(q) / float(p))))
    except Exception:
        pass
    return max(1, n)


def make_frames(n, first_seed=0):
    from sassd_b200.synth import synth_cloud
    return [synth_cloud(first_seed + i) for i in range(n)]


# ------------------------------------------------------------------------------------------- reference arm
REF_TIME_BOX_S = 90.0


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (numba voxelizer / spconv CPU / torch CPU convs /
    CPU NMS), restated in oracle/ (the reference's own packages do not import here: spconv v1.0 and
    mmcv are absent, iou3d is CUDA-only).  All host threads; each step = one frame."""
    if rank != 0:
        return
    from oracle import ref_pipeline as O
    from sassd_b200.checkpoint import make_synthetic_state_dict
    torch.set_num_threads(usable_cores())
    sd = make_synthetic_state_dict(0, 1)
    frames = make_frames(max(2, min(args.steps, 8)))
    for i in range(max(1, args.warmup)):
        O.forward_test(sd, [frames[i % len(frames)]], ORACLE_CFG)
    # one frame per step; the run is time-boxed (~1.2 s per frame on 16 cores): after REF_TIME_BOX_S the remaining
    # steps are not executed and the rate of the frames that were timed is reported (steps_timed says how many)
    t0 = time.perf_counter()
    ndet, done = 0, 0
    for i in range(args.steps):
        det = O.forward_test(sd, [frames[i % len(frames)]], ORACLE_CFG)
        ndet += 0 if det[0][0] is None else len(det[0][0])
        done += 1
        if time.perf_counter() - t0 > REF_TIME_BOX_S:
            break
    dt = time.perf_counter() - t0
    fps = done / dt
    cores = torch.get_num_threads()
    line = dict(metric=METRIC, value=fps, unit="frames/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * dt / done, steps_timed=done, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload="car_cfg.py single-class inference, batch=1, synthetic HDL-64E clouds (~20k pts)",
                            frames_per_step=1),
                cpu_baseline=dict(value=fps, unit="frames/s", cores=cores, kind="port",
                                  sample="%d frames (one per step), CPU oracle port of the reference path" % done),
                e2e=dict(value=fps, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                detections=ndet)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- our arm
def algorithmic_work(aux, batch):
    """Pair-model bytes / flops of the ruled sparse convs of one step from the rulebooks."""
    plan = [("subm0", [(4, 16), (16, 16)]), ("down0", [(16, 32)]), ("subm1", [(32, 32), (32, 32)]),
            ("down1", [(32, 64)]), ("subm2", [(64, 64)] * 3), ("down2", [(64, 64)]), ("subm3", [(64, 64)] * 3)]
    books = aux["sparse"].indice_dict
    tot_b, tot_f, pairs = 0, 0, {}
    for key, layers in plan:
        rb = books[key]
        n = int(rb.d_rows_out.item())
        p = int((rb.nbr[:n] >= 0).sum().item())
        pairs[key] = p
        for cin, cout in layers:
            tot_b += p * (4 * cin + 4 * cout + 8)
            tot_f += 2 * p * cin * cout
    return tot_b, tot_f, pairs


# DRAM bytes of one B=1 launch of the roofline kernel, from the committed `ncu --set full` capture
NCU_DRAM_BYTES_PER_LAUNCH = {"tma::conv2d_tma_kernel<256>": 38518016 + 2328064}


def profile_step(model, points, pt_off, batch, maxpts, iters=3):
    """Per-C-ABI-call CUDA-event timing of one step (events on the launching stream)."""
    from sassd_b200 import ops
    agg = {}
    for _ in range(iters):
        ops.PROFILE = []
        # hold the GPU for ~3 ms so that the host has queued the whole step before the first kernel starts: the
        # event pairs then bracket back-to-back kernels instead of kernel + host launch latency
        torch.cuda._sleep(6_000_000)
        det, nd, status, aux = model.forward_device(points, pt_off, batch, maxpts)
        torch.cuda.synchronize()
        for name, label, e0, e1 in ops.PROFILE:
            k = label or name
            a = agg.setdefault(k, [0.0, 0])
            a[0] += e0.elapsed_time(e1); a[1] += 1
        ops.PROFILE = None
    return {k: dict(ms_total_per_step=v[0] / iters, calls_per_step=v[1] // iters) for k, v in agg.items()}, aux


def run_ours(args, rank, world, local):
    import sassd_b200 as S
    from sassd_b200 import checkpoint, dist as D, ops
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback in the product path)"
    torch.set_num_threads(min(8, usable_cores()))     # host side only stages buffers; keep the pools small
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    cfg = S.Config.fromfile(os.path.join(ROOT, "configs", "car_cfg.py"))
    model, vg, aset = S.build_from_config(cfg, device=str(dev))
    sd = checkpoint.make_synthetic_state_dict(0, 1)
    checkpoint.load_state_dict_into(model, sd)
    if args.precision == "tf32x3":
        model.set_precision(ops.PREC_TF32X3)
    elif args.precision == "f16x3":
        model.set_precision(ops.PREC_F16X3)
    elif args.precision == "mixed":          # tensor cores for the dense convs, FFMA for the sparse backbone
        model.set_precision(ops.PREC_TF32X3, sparse=ops.PREC_FP32)
    B = args.batch
    pool = 8
    frames = make_frames(pool * B, first_seed=rank * 1000)      # every rank owns its own frames (weak scaling)
    batches = [frames[i * B:(i + 1) * B] for i in range(pool)]
    # device-resident copies for the kernel-side number
    staged = []
    for fb in batches:
        hp, ho, counts = model.stage_points(fb)
        staged.append((hp.to(dev).clone(), ho.to(dev).clone(), max(counts)))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    graph = None
    if not args.no_graph:
        maxpts = ops.next_pow2(max(max(p.shape[0] for p in fb) for fb in batches))
        graph = model.enable_cuda_graph(B, maxpts)

    def step(i):
        p, o, mx = staged[i % pool]
        if graph is not None:
            graph.load_device(p, o)
            return graph.replay() + (None,)
        return model.forward_device(p, o, B, mx)

    for i in range(max(3, args.warmup)):
        det, nd, status, aux = step(i)
    torch.cuda.synchronize()
    word = int(status.item())
    assert word == 0, "device status flags %s" % ops._lib.decode_flags(word)

    # ---- timed region: K steps, CUDA events on the launching stream, L2 flushed (untimed) between steps
    sampler = ClockSampler(local)
    D.barrier(); torch.cuda.synchronize()
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = ops.LAUNCHES
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()
        evs[i][0].record()
        det, nd, status, aux = step(i)
        evs[i][1].record()
    torch.cuda.synchronize()
    launches = (ops.LAUNCHES - l0)
    if graph is not None:      # launches are inside the captured graph: count the kernels of one eager step
        l1 = ops.LAUNCHES
        model.forward_device(*staged[0][:2], B, staged[0][2])
        torch.cuda.synchronize()
        launches = (ops.LAUNCHES - l1) * args.steps
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    # the shard's single exchange step: gather of the fixed-size results (NCCL when world > 1)
    g0 = torch.cuda.Event(enable_timing=True); g1 = torch.cuda.Event(enable_timing=True)
    g0.record()
    det_all, nd_all = D.gather_detections(det, nd)
    g1.record()
    torch.cuda.synchronize()
    dev_ms += g0.elapsed_time(g1)
    D.barrier(); torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    dev_ms = D.max_over_ranks(dev_ms, dev)
    value = world * args.steps * B / (dev_ms / 1e3)

    # ---- e2e through the public API: host numpy points -> pinned -> H2D -> path -> D2H detections
    # throughput API: detect_stream (double-buffered CUDA graphs; H2D of step i+1 overlaps the GPU work of step i)
    maxpts_e2e = ops.next_pow2(max(max(p.shape[0] for p in fb) for fb in batches))
    for _ in model.detect_stream([batches[i % pool] for i in range(2 * args.in_flight)], B, maxpts_e2e,
                                 depth=args.in_flight, concurrent=not args.serial_stream):
        pass
    D.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    ndet = 0
    for out in model.detect_stream((batches[i % pool] for i in range(args.steps)), B, maxpts_e2e,
                                   depth=args.in_flight, concurrent=not args.serial_stream):
        ndet += sum(0 if o["boxes_lidar"] is None else len(o["boxes_lidar"]) for o in out)
    torch.cuda.synchronize()
    e2e_s = D.max_over_ranks(time.perf_counter() - t0, dev)
    e2e = world * args.steps * B / e2e_s
    # the same stream with one step on the GPU at a time, for comparison
    t0 = time.perf_counter()
    for out in model.detect_stream((batches[i % pool] for i in range(args.steps)), B, maxpts_e2e,
                                   depth=args.in_flight, concurrent=False):
        pass
    torch.cuda.synchronize()
    e2e_serial = world * args.steps * B / D.max_over_ranks(time.perf_counter() - t0, dev)
    if os.environ.get("SASSD_BENCH_DEPTHS") and rank == 0:      # experiment: other numbers of steps in flight
        for dpt in [int(v) for v in os.environ["SASSD_BENCH_DEPTHS"].split(",")]:
            for _ in model.detect_stream([batches[i % pool] for i in range(2 * dpt)], B, maxpts_e2e, depth=dpt):
                pass
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for out in model.detect_stream((batches[i % pool] for i in range(args.steps)), B, maxpts_e2e, depth=dpt):
                pass
            torch.cuda.synchronize()
            print("in-flight %d: %.1f frames/s" % (dpt, args.steps * B / (time.perf_counter() - t0)), file=sys.stderr)
    # latency of the synchronous single call (stage + H2D + graph + D2H + sync), for reference
    t0 = time.perf_counter()
    for i in range(min(args.steps, 10)):
        model.forward_points(batches[i % pool])
    sync_ms = 1e3 * (time.perf_counter() - t0) / min(args.steps, 10)
    h2d = int(np.mean([sum(p.shape[0] for p in fb) * 16 + (B + 1) * 4 for fb in batches]))
    d2h = int(det.numel() * 4 + nd.numel() * 4 + 4)

    if rank != 0:
        return
    # ---- per-kernel profile (rank 0): dominant kernel + sparse-conv roofline
    prof, aux = profile_step(model, *staged[0][:2], B, staged[0][2])
    peaks = load_peaks()
    H, W = 200, 176
    dom = max(prof.items(), key=lambda kv: kv[1]["ms_total_per_step"])
    bev_key = "conv2d_tma[taps=9 256->256]" if "conv2d_tma[taps=9 256->256]" in prof else "gconv[conv2d taps=9 256->256]"
    bev = prof.get(bev_key)
    roofline = None
    if bev:
        # Constant-region tile skipping: only the tiles that are actually computed count as work
        tiles_frac = 1.0
        dist = getattr(aux.get("x"), "tile_dist", None)
        if dist is not None:
            from sassd_b200.lib import CONV2D_TILE_H as TH, CONV2D_TILE_W as TW
            ty, tx = (H + TH - 1) // TH, (W + TW - 1) // TW
            d = dist.cpu().numpy().reshape(B, ty, tx)
            border = np.zeros((ty, tx), bool)
            border[0] = border[-1] = True
            border[:, 0] = border[:, -1] = True
            # the six 3x3 256->256 layers are 2..7 convolutions away from the scattered map
            tiles_frac = float(np.mean([((d <= reach) | border[None]).mean() for reach in range(2, 8)]))
        flops = 2.0 * B * H * W * 9 * 256 * 256 * tiles_frac
        per_launch_ms = bev["ms_total_per_step"] / bev["calls_per_step"]
        ach = flops / (per_launch_ms * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        kname = {"fp32": "gconv_ffma_kernel<CONV2D,128,16>", "tf32x3": "tc::gconv_tc_kernel<CONV2D,256,1,TF32X3>",
                 "mixed": "tc::gconv_tc_kernel<CONV2D,256,1,TF32X3>",
                 "f16x3": "tma::conv2d_tma_kernel<256>" if bev_key.startswith("conv2d_tma") else
                          "tc::gconv_tc_kernel<CONV2D,256,1,F16X3>"}[args.precision]
        passes = {"fp32": "fp32 FFMA (no tensor cores)", "tf32x3": "3 TF32 MMA passes per algorithmic flop",
                  "mixed": "3 TF32 MMA passes per algorithmic flop",
                  "f16x3": "3 FP16 MMA passes per algorithmic flop (ceiling 1/3 of the fp16/bf16 peak)"}[args.precision]
        roofline = dict(kernel="%s (BEVNet 3x3 256->256, %d launches/step)" % (kname, bev["calls_per_step"]),
                        bound="tensor", achieved=ach, peak=peak, unit="TFLOP/s", frac=ach / peak,
                        traffic=(NCU_DRAM_BYTES_PER_LAUNCH.get(kname) and NCU_DRAM_BYTES_PER_LAUNCH[kname] * B)
                        if tiles_frac == 1.0 else None,
                        traffic_all_tiles=NCU_DRAM_BYTES_PER_LAUNCH.get(kname) and NCU_DRAM_BYTES_PER_LAUNCH[kname] * B,
                        traffic_unit="bytes per launch, dram__bytes_read.sum + dram__bytes_write.sum of the B=1 launch "
                                     "in profiles/r1_ncu_full_conv2d_tma.md (every tile computed), scaled by the batch",
                        tiles_computed_frac=tiles_frac,
                        peak_source="%s bf16 dense, sustained" % peaks["source"],
                        note="achieved = algorithmic fp32 flops of the COMPUTED tiles / CUDA-event time (tiles in the "
                             "map's constant region are stored, not computed: tiles_computed_frac); " + passes,
                        mma_issue_frac=(3.0 if args.precision != "fp32" else 1.0) * ach / peak *
                                       (2.0 if args.precision in ("tf32x3", "mixed") else 1.0),
                        share_of_step=bev["ms_total_per_step"] / sum(v["ms_total_per_step"] for v in prof.values()))
    sp_bytes, sp_flops, pairs = algorithmic_work(aux, B)
    sp_ms = sum(v["ms_total_per_step"] for k, v in prof.items() if k.startswith("gconv[table") or k.startswith("spconv_split[taps=27"))
    sp_kernel = {"fp32": "gconv_ffma_kernel<TABLE,...>", "mixed": "gconv_ffma_kernel<TABLE,...>"}.get(
        args.precision, "tc::gconv_tc_kernel<TABLE,BN,1,%s>" % args.precision.upper())
    if any(k.startswith("spconv_split") for k in prof):
        sp_kernel = "sps::spconv_split_kernel<TABLE,BN> (async cp.async gather of split fp16 rows, tap-packed tcgen05 FP16x3)"
    sp_ach = sp_bytes / (sp_ms * 1e-3) / 1e9 if sp_ms > 0 else 0.0
    roofline_sparse = dict(kernel=sp_kernel + " x13 ruled sparse convs", bound="hbm", achieved=sp_ach,
                           peak=peaks["hbm_gbs"], unit="GB/s", frac=sp_ach / peaks["hbm_gbs"], traffic=None,
                           algorithmic_bytes_per_step=sp_bytes, flops_per_step=sp_flops, ms_per_step=sp_ms,
                           pairs=pairs, peak_source=peaks["source"])
    stages = {}
    for k, v in prof.items():
        stages[k] = round(v["ms_total_per_step"], 4)

    # ---- CPU baseline (oracle port) on a bounded sample, rank 0 at N=1 only
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import ref_pipeline as O
        torch.set_num_threads(usable_cores())
        O.forward_test(sd, [frames[0]], ORACLE_CFG)
        nsamp = 4
        t0 = time.perf_counter()
        for i in range(nsamp):
            O.forward_test(sd, [frames[i % len(frames)]], ORACLE_CFG)
        dt = time.perf_counter() - t0
        cpu = dict(value=nsamp / dt, unit="frames/s", cores=torch.get_num_threads(), kind="port",
                   sample="%d frames of the same workload through the CPU oracle (C voxelizer, torch-CPU "
                          "gather/mm/scatter sparse conv + conv2d, C rotated NMS)" % nsamp)

    line = dict(metric=METRIC, value=value, unit="frames/s", n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
                ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="ours",
                config=dict(workload="car_cfg.py single-class inference, batch=%d, synthetic HDL-64E clouds (~20k pts), "
                                     "raw points -> detections" % B,
                            frames_per_step=B, weights="synthetic (seed 0, BN calibrated)",
                            l2="flushed between steps (256 MiB memset, untimed)", precision=args.precision,
                            bev_tile_skipping=bool(ops.TILE_OCCUPANCY),
                            cuda_graph=graph is not None,
                            parallelism="frames sharded, dp%d" % world),
                clocks=clocks, gpu_launches=launches,
                e2e=dict(value=e2e, unit="frames/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         api="SingleStageDetector.detect_stream (host numpy points in, numpy detections out)",
                         steps_in_flight=1 if args.serial_stream else args.in_flight, value_one_step_in_flight=e2e_serial,
                         sync_call_ms=sync_ms),
                roofline=roofline, roofline_sparse=roofline_sparse, cpu_baseline=cpu,
                stages_ms=stages, dominant=dom[0], wall_s=t_wall, detections_e2e=ndet)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="f16x3", choices=["fp32", "tf32x3", "f16x3", "mixed"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of one CUDA graph")
    ap.add_argument("--serial-stream", action="store_true",
                    help="e2e: one step on the GPU at a time (default: detect_stream keeps --in-flight captured steps going)")
    ap.add_argument("--in-flight", type=int, default=4, help="captured steps detect_stream keeps in flight (e2e)")
    args = ap.parse_args()
    from sassd_b200 import dist as D
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference(args, rank, int(os.environ.get("WORLD_SIZE", "1")))
        return
    rank, world, local = D.init_from_env()
    try:
        run_ours(args, rank, world, local)
    finally:
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
