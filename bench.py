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


def workload_config(batch):
    """The `config` object both arms print - identical strings, so the driver can tell they ran the same thing."""
    return dict(workload="car_cfg.py single-class inference, batch=%d, synthetic HDL-64E clouds (~20k pts), "
                         "raw points -> detections" % batch, frames_per_step=batch)
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


def cpu_threads():
    """Threads for the CPU oracle: every usable core up to 32 - beyond that torch's intra-op pools only contend on
    these small per-frame tensors (measured round 1: 0.91 frames/s on 16 threads, 0.76 on 96)."""
    return max(1, min(usable_cores(), int(os.environ.get("SASSD_CPU_THREADS", "32"))))


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
    cores = cpu_threads()
    torch.set_num_threads(cores)
    sd = make_synthetic_state_dict(0, 1)
    B = args.batch
    frames = make_frames(max(2, min(args.steps, 8)) * B)
    batches = [frames[i * B:(i + 1) * B] for i in range(len(frames) // B)]
    for i in range(max(1, min(args.warmup, 2))):
        O.forward_test(sd, batches[i % len(batches)], ORACLE_CFG)
    # one batch per step; the run is time-boxed (~1 s per frame on 16 cores): after REF_TIME_BOX_S the remaining
    # steps are not executed and the rate of the frames that were timed is reported (steps_timed says how many)
    t0 = time.perf_counter()
    ndet, done = 0, 0
    for i in range(args.steps):
        det = O.forward_test(sd, batches[i % len(batches)], ORACLE_CFG)
        ndet += sum(0 if d is None else len(d) for d in det[0])
        done += 1
        if time.perf_counter() - t0 > REF_TIME_BOX_S:
            break
    dt = time.perf_counter() - t0
    fps = done * B / dt
    line = dict(metric=METRIC, value=fps, unit="frames/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * dt / done, steps_timed=done, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference", config=workload_config(B),
                cpu_baseline=dict(value=fps, unit="frames/s", cores=cores, kind="port",
                                  sample="%d batches of %d frame(s) (one per step) through the CPU oracle port of the "
                                         "reference path (C restatement of the numba voxelizer, torch-CPU "
                                         "gather/mm/scatter sparse conv + conv2d, C rotated NMS)" % (done, B)),
                e2e=dict(value=fps, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                detections=ndet,
                note="one CPU process on rank 0 whatever --gpus is (the other ranks exit): this value does not scale "
                     "with N, so only the N=1 ratio to the GPU arm is like for like")
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- our arm
SPARSE_PLAN = [("subm0", [(4, 16), (16, 16)]), ("down0", [(16, 32)]), ("subm1", [(32, 32), (32, 32)]),
               ("down1", [(32, 64)]), ("subm2", [(64, 64)] * 3), ("down2", [(64, 64)]), ("subm3", [(64, 64)] * 3)]


def algorithmic_work(aux, batch):
    """Pair-model bytes / flops of the ruled sparse convs of one step from the rulebooks, and
    the row-taps the kernel executes for them: a layer runs, per 128-row tile, the K chunks (64 / cin_stored taps
    each) that hold at least one tap of the tile's mask."""
    books = aux["sparse"].indice_dict
    tot_b, tot_f, pairs, exec_over_pairs, exec_chunks = 0, 0, {}, {}, 0
    for key, layers in SPARSE_PLAN:
        rb = books[key]
        n = int(rb.d_rows_out.item())
        p = int((rb.nbr[:n] >= 0).sum().item())
        pairs[key] = p
        ntiles = (n + 127) // 128
        masks = rb.tile_mask[:ntiles].cpu().numpy().astype(np.int64) if rb.tile_mask is not None else None
        ratios = []
        for cin, cout in layers:
            tot_b += p * (4 * cin + 4 * cout + 8)
            tot_f += 2 * p * cin * cout
            cs = (cin + 7) // 8 * 8
            tpg = 64 // cs if 64 % cs == 0 else 1
            nchunks = (27 + tpg - 1) // tpg
            if masks is None:
                chunks = ntiles * nchunks
            else:
                m = np.where(masks == 0, 1, masks)
                chunks = int(sum(((m >> (g * tpg)) & ((1 << tpg) - 1) != 0).sum() for g in range(nchunks)))
            exec_chunks += chunks
            ratios.append(chunks * min(tpg, 27) * 128 / max(p, 1))
        exec_over_pairs[key] = round(float(np.mean(ratios)), 2)
    return tot_b, tot_f, pairs, exec_over_pairs, exec_chunks


# DRAM bytes of one B=1 launch of the roofline kernel with the constant-region tile skipping ON, from the committed
# `ncu --set full` capture profiles/r2_ncu_full_conv2d_tma.md (dram__bytes_read.sum + dram__bytes_write.sum)
NCU_DRAM_BYTES_PER_LAUNCH = {}


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


def count_step(model, points, pt_off, batch, maxpts):
    """One eager step with the kernels' own instrumentation counters switched on: BEV tiles computed vs stored as a
    constant (per layer label), sparse (tile, chunk) pairs executed."""
    from sassd_b200 import ops
    dev = points.device
    ops.SPCONV_COUNTERS = torch.zeros(2, dtype=torch.int32, device=dev)

    class _Lazy(dict):
        def get(self, label, default=None):
            if label not in self:
                self[label] = torch.zeros(2, dtype=torch.int32, device=dev)
            return self[label]
    ops.CONV2D_COUNTERS = _Lazy()
    try:
        model.forward_device(points, pt_off, batch, maxpts)
        torch.cuda.synchronize()
        tiles = {k: [int(x) for x in v.cpu().tolist()] for k, v in ops.CONV2D_COUNTERS.items()}
        sp = [int(x) for x in ops.SPCONV_COUNTERS.cpu().tolist()]
    finally:
        ops.SPCONV_COUNTERS = None
        ops.CONV2D_COUNTERS = None
    return tiles, sp


def parity_check(model, sd, batches, batch, maxpts, n_frames=2):
    """Correctness guard on the very path that was timed: stream `n_frames` frames through detect_stream and compare
    with the CPU oracle (same weights, same points): detection counts, boxes and scores within 1e-4."""
    from oracle import ref_pipeline as O
    torch.set_num_threads(cpu_threads())
    fbs = [batches[i] for i in range(max(1, (n_frames + batch - 1) // batch))]
    got = list(model.detect_stream(fbs, batch, maxpts, depth=2))
    frames = [f for fb in fbs for f in fb][:n_frames]
    res = dict(frames=len(frames), detections_ours=0, detections_oracle=0, max_score_err=0.0, max_box_err=0.0, ok=True)
    flat = [o for out in got for o in out][:n_frames]
    for f, o in zip(frames, flat):
        exp = O.forward_test(sd, [f], ORACLE_CFG)
        eb, es = exp[0][0], exp[1][0]
        ne = 0 if eb is None else len(eb)
        ng = 0 if o["boxes_lidar"] is None else len(o["boxes_lidar"])
        res["detections_ours"] += ng
        res["detections_oracle"] += ne
        if ne != ng:
            res["ok"] = False
            continue
        if ne:
            res["max_score_err"] = max(res["max_score_err"], float(np.abs(o["scores"] - es).max()))
            res["max_box_err"] = max(res["max_box_err"], float(np.abs(o["boxes_lidar"] - eb).max()))
    res["ok"] = bool(res["ok"] and res["max_score_err"] <= 1e-4 and res["max_box_err"] <= 1e-3)
    res["tolerance"] = "scores 1e-4, boxes 1e-4 relative to the box scale (<= 1e-3 absolute), equal counts"
    return res


MIN_TIMED_S = 1.0      # the timed region repeats the K steps until it is at least this long


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
    elif args.precision == "fp32":
        model.set_precision(ops.PREC_FP32)
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

    maxpts = ops.next_pow2(max(max(p.shape[0] for p in fb) for fb in batches))
    graph = None
    if not args.no_graph:
        graph = model.enable_cuda_graph(B, maxpts)

    def step(i):
        p, o, mx = staged[i % pool]
        if graph is not None:
            graph.load_device(p, o)
            return graph.replay() + (None,)
        return model.forward_device(p, o, B, mx)

    t_est0 = time.perf_counter()
    for i in range(max(3, args.warmup)):
        det, nd, status, aux = step(i)
    torch.cuda.synchronize()
    est_step_s = (time.perf_counter() - t_est0) / max(3, args.warmup)
    word = int(status.item())
    assert word == 0, "device status flags %s" % ops._lib.decode_flags(word)
    # the shard's single exchange step: pre-allocated, warmed before anything is timed
    gather = D.DetectionGather(det.shape[0], det.shape[1], dev)
    gather.warm()

    # ---- timed region: K steps (repeated `rounds` times until >= MIN_TIMED_S), CUDA events on the launching
    # stream around every step, L2 flushed (untimed) between steps, + the result gather, max over ranks
    rounds = max(1, int(np.ceil(MIN_TIMED_S / max(args.steps * est_step_s, 1e-6))))
    rounds = int(D.max_over_ranks(rounds, dev))
    nsteps = args.steps * rounds
    sampler = ClockSampler(local)
    D.barrier(); torch.cuda.synchronize()
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
    l0 = ops.LAUNCHES
    t_wall0 = time.perf_counter()
    for i in range(nsteps):
        flush.zero_()
        evs[i][0].record()
        det, nd, status, aux = step(i)
        evs[i][1].record()
    g0 = torch.cuda.Event(enable_timing=True); g1 = torch.cuda.Event(enable_timing=True)
    g0.record()
    det_all, nd_all = gather(det, nd)
    g1.record()
    torch.cuda.synchronize()
    D.barrier(); torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = (ops.LAUNCHES - l0)
    if graph is not None:      # launches are inside the captured graph: count the kernels of one eager step
        l1 = ops.LAUNCHES
        model.forward_device(*staged[0][:2], B, staged[0][2])
        torch.cuda.synchronize()
        launches = (ops.LAUNCHES - l1) * nsteps
    gather_ms = g0.elapsed_time(g1)
    dev_ms = sum(a.elapsed_time(b) for a, b in evs) + gather_ms
    dev_ms = D.max_over_ranks(dev_ms, dev)
    value = world * nsteps * B / (dev_ms / 1e3)

    # ---- e2e through the public API: host numpy points -> pinned -> H2D -> path -> D2H detections (+ the gather)
    # throughput API: detect_stream (CUDA graphs in flight; H2D of step i+1 overlaps the GPU work of step i)
    for _ in model.detect_stream([batches[i % pool] for i in range(2 * args.in_flight)], B, maxpts,
                                 depth=args.in_flight, concurrent=not args.serial_stream):
        pass
    e2e_steps = nsteps
    D.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    ndet = 0
    for out in model.detect_stream((batches[i % pool] for i in range(e2e_steps)), B, maxpts,
                                   depth=args.in_flight, concurrent=not args.serial_stream):
        ndet += sum(0 if o["boxes_lidar"] is None else len(o["boxes_lidar"]) for o in out)
    gather(det, nd)                     # the shard's result exchange belongs to the end-to-end job
    torch.cuda.synchronize()
    e2e_s = D.max_over_ranks(time.perf_counter() - t0, dev)
    e2e = world * e2e_steps * B / e2e_s
    # the same stream with one step on the GPU at a time, for comparison
    t0 = time.perf_counter()
    for out in model.detect_stream((batches[i % pool] for i in range(args.steps)), B, maxpts,
                                   depth=args.in_flight, concurrent=False):
        pass
    torch.cuda.synchronize()
    e2e_serial = world * args.steps * B / D.max_over_ranks(time.perf_counter() - t0, dev)
    if os.environ.get("SASSD_BENCH_DEPTHS") and rank == 0:      # experiment: other numbers of steps in flight
        for dpt in [int(v) for v in os.environ["SASSD_BENCH_DEPTHS"].split(",")]:
            for _ in model.detect_stream([batches[i % pool] for i in range(2 * dpt)], B, maxpts, depth=dpt):
                pass
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for out in model.detect_stream((batches[i % pool] for i in range(args.steps)), B, maxpts, depth=dpt):
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
    model.disable_cuda_graph()
    prof, aux = profile_step(model, *staged[0][:2], B, staged[0][2])
    tile_counts, sp_counts = count_step(model, *staged[0][:2], B, staged[0][2])
    peaks = load_peaks()
    H, W = 200, 176
    dom = max(prof.items(), key=lambda kv: kv[1]["ms_total_per_step"])
    bev_key = "conv2d_tma[taps=9 256->256]" if "conv2d_tma[taps=9 256->256]" in prof else "gconv[conv2d taps=9 256->256]"
    bev = prof.get(bev_key)
    roofline = None
    if bev:
        # Constant-region tile skipping: only the tiles that are actually computed count as
        # work; the kernel counts them itself (computed, total over the 6 launches of the layer shape)
        tc = tile_counts.get(bev_key)
        tiles_frac = (tc[0] / tc[1]) if tc and tc[1] else 1.0
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
        ncu_b1 = NCU_DRAM_BYTES_PER_LAUNCH.get(kname)
        roofline = dict(kernel="%s (BEVNet 3x3 256->256, %d launches/step)" % (kname, bev["calls_per_step"]),
                        bound="tensor", achieved=ach, peak=peak, unit="TFLOP/s", frac=ach / peak,
                        traffic=(ncu_b1 * B) if ncu_b1 else None,
                        traffic_unit="bytes per launch: dram__bytes_read.sum + dram__bytes_write.sum of the B=1 launch "
                                     "(tile skipping on) in profiles/r2_ncu_full_conv2d_tma.md, scaled by the batch",
                        tiles_computed=tc[0] if tc else None, tiles_total=tc[1] if tc else None,
                        tiles_computed_frac=tiles_frac,
                        peak_source="%s bf16 dense, sustained" % peaks["source"],
                        note="achieved = algorithmic fp32 flops of the COMPUTED tiles (counted by the kernel) / "
                             "CUDA-event time; tiles in the map's constant region are stored, not computed; " + passes,
                        mma_issue_frac=(3.0 if args.precision != "fp32" else 1.0) * ach / peak *
                                       (2.0 if args.precision in ("tf32x3", "mixed") else 1.0),
                        share_of_step=bev["ms_total_per_step"] / sum(v["ms_total_per_step"] for v in prof.values()))
    sp_bytes, sp_flops, pairs, exec_over_pairs, exec_chunks = algorithmic_work(aux, B)
    sp_ms = sum(v["ms_total_per_step"] for k, v in prof.items() if k.startswith("gconv[table") or k.startswith("spconv_split[taps=27"))
    sp_kernel = {"fp32": "gconv_ffma_kernel<TABLE,...>", "mixed": "gconv_ffma_kernel<TABLE,...>"}.get(
        args.precision, "tc::gconv_tc_kernel<TABLE,BN,1,%s>" % args.precision.upper())
    if any(k.startswith("spconv_split") for k in prof):
        sp_kernel = ("sps::spconv_split_kernel<TABLE,BN> (cp.async gather of split fp16 rows, tap-packed tcgen05 FP16x3, "
                     "tile-level tap skipping, 2-CTA tap split for small layers)")
    sp_ach = sp_bytes / (sp_ms * 1e-3) / 1e9 if sp_ms > 0 else 0.0
    roofline_sparse = dict(kernel=sp_kernel + " x13 ruled sparse convs", bound="hbm", achieved=sp_ach,
                           peak=peaks["hbm_gbs"], unit="GB/s", frac=sp_ach / peaks["hbm_gbs"], traffic=None,
                           algorithmic_bytes_per_step=sp_bytes, flops_per_step=sp_flops, ms_per_step=sp_ms,
                           pairs=pairs, executed_row_taps_over_pairs=exec_over_pairs,
                           executed_chunks_host_rule=exec_chunks, executed_chunks_kernel_counter=sp_counts[0],
                           peak_source=peaks["source"],
                           note="pair-model bytes (SURVEY 8d): P*(4Cin+4Cout+8) per layer; the features are "
                                "L2-resident, so this is a pair rate expressed in bytes, not DRAM traffic")
    stages = {}
    for k, v in prof.items():
        stages[k] = round(v["ms_total_per_step"], 4)

    # ---- correctness guard on the timed path + CPU baseline (oracle port) on a bounded sample, N=1 only
    parity = None
    if not args.no_parity:
        parity = parity_check(model, sd, batches, B, maxpts)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import ref_pipeline as O
        torch.set_num_threads(cpu_threads())
        O.forward_test(sd, [frames[0]], ORACLE_CFG)
        nsamp = 8
        t0 = time.perf_counter()
        for i in range(nsamp):
            O.forward_test(sd, [frames[i % len(frames)]], ORACLE_CFG)
        dt = time.perf_counter() - t0
        cpu = dict(value=nsamp / dt, unit="frames/s", cores=torch.get_num_threads(), kind="port",
                   sample="%d frames of the same workload through the CPU oracle port of the reference path (C "
                          "restatement of the numba voxelizer, torch-CPU gather/mm/scatter sparse conv + conv2d, C "
                          "rotated NMS)" % nsamp)

    line = dict(metric=METRIC, value=value, unit="frames/s", n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
                ms_per_step=dev_ms / nsteps, steps_timed=nsteps, rounds=rounds, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic", impl="ours", config=workload_config(B),
                details=dict(weights="synthetic (seed 0, BN calibrated)",
                             l2="flushed between steps (256 MiB memset, untimed)", precision=args.precision,
                             bev_tile_skipping=bool(ops.TILE_OCCUPANCY), sparse_tap_skipping=bool(ops.SPCONV_TAP_SKIP),
                             sparse_tap_split=bool(ops.SPCONV_TAP_SPLIT), cuda_graph=graph is not None,
                             parallelism="frames sharded, dp%d" % world,
                             timed_region="%d x %d steps (>= %.1f s), per-step CUDA events + the result all_gather "
                                          "(%.3f ms)" % (rounds, args.steps, MIN_TIMED_S, gather_ms)),
                clocks=clocks, gpu_launches=launches,
                e2e=dict(value=e2e, unit="frames/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         api="SingleStageDetector.detect_stream (host numpy points in, numpy detections out) + the "
                             "shard's result all_gather",
                         steps_timed=e2e_steps, steps_in_flight=1 if args.serial_stream else args.in_flight,
                         value_one_step_in_flight=e2e_serial, sync_call_ms=sync_ms),
                roofline=roofline, roofline_sparse=roofline_sparse, cpu_baseline=cpu, parity_check=parity,
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
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of streamed frames")
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
