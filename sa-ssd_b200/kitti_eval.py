"""KITTI official-style evaluation: bbox / BEV / 3-D AP (R11) and AOS.

Mirrors the reference's evaluator so that `tools/test.py`-style code can swap the import:

* `get_official_eval_result(gt_annos, dt_annos, current_classes, difficultys)` -> the same printed table
  (mmdet/core/evaluation/kitti_eval.py:791-851); `official_eval` additionally returns the AP arrays;
* `rotate_iou_gpu_eval(boxes, query_boxes, criterion)` <- mmdet/core/post_processing/rotate_nms_gpu.py:592-627.

What runs where: the rotated-box overlaps of every frame are ONE launch of `sassd_rotate_overlap_eval` (the
reference launches a numba.cuda kernel per 1/50th of the dataset); the 2-D box and height overlaps are vectorised
float64 numpy; the greedy matching + per-threshold statistics (numba-jitted CPU loops in the reference) are the
host C++ function `sassd_kitti_match`.  Annotations are dicts of numpy arrays as produced by
`results.kitti_bbox2results` / the reference's `get_label_annos` (camera-frame boxes, dimensions l, h, w).
"""
import ctypes

import numpy as np

from . import lib as _lib

CLASS_NAMES = ['car', 'pedestrian', 'cyclist', 'van', 'person_sitting', 'car', 'tractor', 'trailer']
CLASS_TO_NAME = {0: 'Car', 1: 'Pedestrian', 2: 'Cyclist', 3: 'Van', 4: 'Person_sitting', 5: 'car', 6: 'tractor',
                 7: 'trailer'}
MIN_HEIGHT, MAX_OCCLUSION, MAX_TRUNCATION = (40, 25, 25), (0, 1, 2), (0.15, 0.3, 0.5)
# [metric (bbox, bev, 3d), class]: the two official overlap sets (kitti_eval.py:792-797)
OVERLAP_0_7 = np.array([[0.7, 0.5, 0.5, 0.7, 0.5, 0.7, 0.7, 0.7]] * 3)
OVERLAP_0_5 = np.array([[0.7, 0.5, 0.5, 0.7, 0.5, 0.5, 0.5, 0.5], [0.5, 0.25, 0.25, 0.5, 0.25, 0.5, 0.5, 0.5],
                        [0.5, 0.25, 0.25, 0.5, 0.25, 0.5, 0.5, 0.5]])


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def _offsets(counts, dtype=np.int32):
    return np.concatenate([[0], np.cumsum(counts)]).astype(dtype)


# ------------------------------------------------------------------ overlaps
def _rotated_overlaps_device(boxes, box_off, query, query_off, out_off, criterion):
    """All frames in one kernel launch (float32, [sum nb*nq] flat)."""
    import torch
    from . import ops
    ops.require_cuda()
    dev = torch.device("cuda")
    total = int(out_off[-1])
    out = torch.zeros((max(total, 1),), dtype=torch.float32, device=dev)
    if total:
        t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (boxes, box_off, query, query_off, out_off)]
        pairs = np.diff(box_off).astype(np.int64) * np.diff(query_off).astype(np.int64)
        ops._call("sassd_rotate_overlap_eval", None, ops._ptr(t[0]), ops._ptr(t[1]), ops._ptr(t[2]), ops._ptr(t[3]),
                  ops._ptr(t[4]), len(box_off) - 1, int(criterion), int(pairs.max()), ops._ptr(out), ops._stream())
    return out[:total].cpu().numpy()


def _rotated_overlaps(boxes_list, query_list, criterion, overlap_fn=None):
    """Per-frame [nb, nq] float32 overlap matrices.  `overlap_fn(boxes, query, criterion)` replaces the CUDA kernel
    (the tests inject the CPU oracle there)."""
    if overlap_fn is not None:
        return [np.asarray(overlap_fn(b.astype(np.float32), q.astype(np.float32), criterion), np.float32)
                .reshape(b.shape[0], q.shape[0]) for b, q in zip(boxes_list, query_list)]
    nb = [b.shape[0] for b in boxes_list]
    nq = [q.shape[0] for q in query_list]
    box_off, query_off = _offsets(nb), _offsets(nq)
    out_off = _offsets(np.asarray(nb, np.int64) * np.asarray(nq, np.int64), np.int64)
    boxes = np.concatenate(boxes_list, 0).astype(np.float32).reshape(-1, 5)
    query = np.concatenate(query_list, 0).astype(np.float32).reshape(-1, 5)
    flat = _rotated_overlaps_device(boxes, box_off, query, query_off, out_off, criterion)
    return [flat[out_off[f]:out_off[f + 1]].reshape(nb[f], nq[f]) for f in range(len(nb))]


def rotate_iou_gpu_eval(boxes, query_boxes, criterion=-1, device_id=0):
    """Reference signature (rotate_nms_gpu.py:592): [N,5] x [K,5] (x, y, dx, dy, angle) -> [N,K] float32 (the
    reference casts to the dtype of its float32 working copy, so float64 inputs come back as float32 too)."""
    boxes, query_boxes = np.asarray(boxes), np.asarray(query_boxes)
    if boxes.shape[0] == 0 or query_boxes.shape[0] == 0:
        return np.zeros((boxes.shape[0], query_boxes.shape[0]), np.float32)
    return _rotated_overlaps([boxes], [query_boxes], criterion)[0]


def image_box_overlap(boxes, query_boxes, criterion=-1):
    """Axis-aligned 2-D overlap, kitti_eval.py:95-122, vectorised (same expressions per element)."""
    b, q = boxes[:, None, :], query_boxes[None, :, :]
    iw = np.minimum(b[..., 2], q[..., 2]) - np.maximum(b[..., 0], q[..., 0])
    ih = np.minimum(b[..., 3], q[..., 3]) - np.maximum(b[..., 1], q[..., 1])
    area_b = (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    area_q = (q[..., 2] - q[..., 0]) * (q[..., 3] - q[..., 1])
    if criterion == -1:
        ua = area_b + area_q - iw * ih
    elif criterion == 0:
        ua = area_b + 0 * area_q
    elif criterion == 1:
        ua = area_q + 0 * area_b
    else:
        ua = np.ones_like(iw)
    hit = (iw > 0) & (ih > 0)
    out = np.zeros(iw.shape, dtype=boxes.dtype)
    out[hit] = (iw * ih)[hit] / ua[hit]
    return out


def _d3_from_bev(boxes, qboxes, rinc, criterion=-1):
    """Height overlap on top of the BEV intersection area, camera frame (kitti_eval.py:130-154)."""
    # rinc stays float32 (rotate_iou_gpu_eval returns its float32 working copy's dtype) and the quotient is stored
    # back into it, i.e. rounded to float32, exactly as d3_box_overlap_kernel does
    rinc = np.asarray(rinc, np.float32)
    iw = (np.minimum(boxes[:, None, 1], qboxes[None, :, 1])
          - np.maximum(boxes[:, None, 1] - boxes[:, None, 4], qboxes[None, :, 1] - qboxes[None, :, 4]))
    a1 = (boxes[:, 3] * boxes[:, 4] * boxes[:, 5])[:, None]
    a2 = (qboxes[:, 3] * qboxes[:, 4] * qboxes[:, 5])[None, :]
    inc = iw * rinc
    if criterion == -1:
        ua = a1 + a2 - inc
    elif criterion == 0:
        ua = a1 + 0 * a2
    elif criterion == 1:
        ua = a2 + 0 * a1
    else:
        ua = np.ones_like(inc)
    out = np.zeros(rinc.shape, np.float32)
    hit = (rinc > 0) & (iw > 0)
    out[hit] = (inc[hit] / ua[hit]).astype(np.float32)
    return out


def calculate_overlaps(gt_annos, dt_annos, metric, overlap_fn=None):
    """overlaps[f][detection, ground truth] as float64 (what calculate_iou_partly(dt, gt, metric) returns)."""
    if metric == 0:
        return [image_box_overlap(np.asarray(d["bbox"], np.float64).reshape(-1, 4),
                                  np.asarray(g["bbox"], np.float64).reshape(-1, 4)) for g, d in zip(gt_annos, dt_annos)]

    def cam(a):
        return np.concatenate([np.asarray(a["location"], np.float64).reshape(-1, 3),
                               np.asarray(a["dimensions"], np.float64).reshape(-1, 3),
                               np.asarray(a["rotation_y"], np.float64).reshape(-1, 1)], 1)
    dts, gts = [cam(d) for d in dt_annos], [cam(g) for g in gt_annos]
    bev = [0, 2, 3, 5, 6]
    if metric == 1:
        ov = _rotated_overlaps([d[:, bev] for d in dts], [g[:, bev] for g in gts], -1, overlap_fn)
        return [o.astype(np.float64) for o in ov]
    rinc = _rotated_overlaps([d[:, bev] for d in dts], [g[:, bev] for g in gts], 2, overlap_fn)
    return [_d3_from_bev(d, g, r).astype(np.float64) for d, g, r in zip(dts, gts, rinc)]


# ------------------------------------------------------------------ matching
def clean_data(gt_anno, dt_anno, current_class, difficulty):
    """Ignore flags of one frame (kitti_eval.py:39-92): 0 evaluate, 1 ignore, -1 other class."""
    cls = CLASS_NAMES[current_class].lower()
    names = np.char.lower(np.asarray(gt_anno["name"], dtype=str)) if len(gt_anno["name"]) else np.zeros((0,), str)
    bbox = np.asarray(gt_anno["bbox"], np.float64).reshape(-1, 4)
    valid = np.where(names == cls, 1, -1)
    if cls == "pedestrian":
        valid = np.where(names == "person_sitting", 0, valid)
    elif cls == "car":
        valid = np.where(names == "van", 0, valid)
    ignore = ((np.asarray(gt_anno["occluded"]) > MAX_OCCLUSION[difficulty])
              | (np.asarray(gt_anno["truncated"]) > MAX_TRUNCATION[difficulty])
              | ((bbox[:, 3] - bbox[:, 1]) <= MIN_HEIGHT[difficulty]))
    ign_gt = np.where((valid == 1) & ~ignore, 0, np.where((valid == 0) | (ignore & (valid == 1)), 1, -1)).astype(np.int32)
    dc = bbox[np.asarray(gt_anno["name"], dtype=str) == "DontCare"] if len(names) else np.zeros((0, 4))
    dnames = np.char.lower(np.asarray(dt_anno["name"], dtype=str)) if len(dt_anno["name"]) else np.zeros((0,), str)
    dbox = np.asarray(dt_anno["bbox"], np.float64).reshape(-1, 4)
    height = np.abs(dbox[:, 3] - dbox[:, 1])
    ign_dt = np.where(height < MIN_HEIGHT[difficulty], 1, np.where(dnames == cls, 0, -1)).astype(np.int32)
    return int((ign_gt == 0).sum()), ign_gt, ign_dt, dc


def get_thresholds(scores, num_gt, num_sample_pts=41):
    """Score thresholds at the 41 sampled recall positions (kitti_eval.py:17-36)."""
    scores = np.sort(np.asarray(scores, np.float64))[::-1]
    current_recall, thresholds = 0.0, []
    for i, score in enumerate(scores):
        l_recall = (i + 1) / num_gt
        r_recall = (i + 2) / num_gt if i < len(scores) - 1 else l_recall
        if (r_recall - current_recall) < (current_recall - l_recall) and i < len(scores) - 1:
            continue
        thresholds.append(score)
        current_recall += 1 / (num_sample_pts - 1.0)
    return np.asarray(thresholds, np.float64)


class _Packed:
    """Per-(class, difficulty) flat arrays for sassd_kitti_match."""

    def __init__(self, gt_annos, dt_annos, overlaps, current_class, difficulty):
        prep = [clean_data(g, d, current_class, difficulty) for g, d in zip(gt_annos, dt_annos)]
        self.nframes = len(gt_annos)
        self.total_valid = sum(p[0] for p in prep)
        ng = [len(p[1]) for p in prep]
        nd = [len(p[2]) for p in prep]
        self.gt_off, self.dt_off = _offsets(ng), _offsets(nd)
        self.dc_off = _offsets([p[3].shape[0] for p in prep])
        self.ov_off = _offsets(np.asarray(ng, np.int64) * np.asarray(nd, np.int64), np.int64)
        cat = lambda xs, shape, dt: (np.ascontiguousarray(np.concatenate(xs, 0), dtype=dt) if len(xs) else       # noqa: E731
                                     np.zeros(shape, dt))
        self.overlaps = cat([np.asarray(o, np.float64).reshape(-1) for o in overlaps], (0,), np.float64)
        self.gt_alpha = cat([np.asarray(g["alpha"], np.float64).reshape(-1) for g in gt_annos], (0,), np.float64)
        self.dt_alpha = cat([np.asarray(d["alpha"], np.float64).reshape(-1) for d in dt_annos], (0,), np.float64)
        self.dt_score = cat([np.asarray(d["score"], np.float64).reshape(-1) for d in dt_annos], (0,), np.float64)
        self.dt_bbox = cat([np.asarray(d["bbox"], np.float64).reshape(-1, 4) for d in dt_annos], (0, 4), np.float64)
        self.dc_bbox = cat([p[3].reshape(-1, 4) for p in prep], (0, 4), np.float64)
        self.ign_gt = cat([p[1] for p in prep], (0,), np.int32)
        self.ign_dt = cat([p[2] for p in prep], (0,), np.int32)

    def match(self, metric, min_overlap, thresholds=None, compute_aos=False):
        L = _lib.load()
        nth = 0 if thresholds is None else len(thresholds)
        pr = np.zeros((max(nth, 1), 4), np.float64)
        tp_scores = np.zeros((max(len(self.gt_alpha), 1),), np.float64)
        ntp = np.zeros((1,), np.int64)
        th = np.ascontiguousarray(thresholds, np.float64) if nth else None
        _lib.check(L.sassd_kitti_match(self.nframes, _p(self.overlaps), _p(self.ov_off), _p(self.gt_off), _p(self.dt_off),
                                       _p(self.dc_off), _p(self.gt_alpha), _p(self.dt_alpha), _p(self.dt_score),
                                       _p(self.dt_bbox), _p(self.dc_bbox), _p(self.ign_gt), _p(self.ign_dt), int(metric),
                                       float(min_overlap), int(bool(compute_aos)), nth, _p(th), _p(pr), _p(tp_scores),
                                       _p(ntp)), "sassd_kitti_match")
        return pr[:nth], tp_scores[:int(ntp[0])]


def eval_class(gt_annos, dt_annos, current_classes, difficultys, metric, min_overlaps, compute_aos=False,
               overlap_fn=None):
    """precision / recall / orientation [class, difficulty, min_overlap, 41] (eval_class_v3, kitti_eval.py:549-657).
    min_overlaps: [num_minoverlap, metric, class]."""
    assert len(gt_annos) == len(dt_annos)
    overlaps = calculate_overlaps(gt_annos, dt_annos, metric, overlap_fn)
    shape = (len(current_classes), len(difficultys), len(min_overlaps), 41)
    precision, recall, aos = np.zeros(shape), np.zeros(shape), np.zeros(shape)
    for m, cls in enumerate(current_classes):
        for l, diff in enumerate(difficultys):
            packed = _Packed(gt_annos, dt_annos, overlaps, cls, diff)
            for k, min_overlap in enumerate(min_overlaps[:, metric, m]):
                _, tp_scores = packed.match(metric, min_overlap)
                thresholds = get_thresholds(tp_scores, packed.total_valid)
                pr, _ = packed.match(metric, min_overlap, thresholds, compute_aos)
                n = len(thresholds)
                with np.errstate(divide="ignore", invalid="ignore"):
                    recall[m, l, k, :n] = pr[:, 0] / (pr[:, 0] + pr[:, 2])
                    precision[m, l, k, :n] = pr[:, 0] / (pr[:, 0] + pr[:, 1])
                    if compute_aos:
                        aos[m, l, k, :n] = pr[:, 3] / (pr[:, 0] + pr[:, 1])
                for i in range(n):      # monotone envelope over the sampled points (:645-651)
                    precision[m, l, k, i] = np.max(precision[m, l, k, i:])
                    recall[m, l, k, i] = np.max(recall[m, l, k, i:])
                    if compute_aos:
                        aos[m, l, k, i] = np.max(aos[m, l, k, i:])
    return dict(recall=recall, precision=precision, orientation=aos)


def get_mAP(prec):
    """11-point interpolated AP over the 41 samples (get_mAP_v2, kitti_eval.py:683-688)."""
    return sum(prec[..., i] for i in range(0, prec.shape[-1], 4)) / 11 * 100


def official_eval(gt_annos, dt_annos, current_classes, difficultys=(0, 1, 2), overlap_fn=None):
    """-> (text, dict(bbox, bev, d3, aos) of AP arrays [class, difficulty, min_overlap])."""
    if not isinstance(current_classes, (list, tuple)):
        current_classes = [current_classes]
    name_to_class = {v: k for k, v in CLASS_TO_NAME.items()}
    classes = [name_to_class[c] if isinstance(c, str) else c for c in current_classes]
    min_overlaps = np.stack([OVERLAP_0_7, OVERLAP_0_5], axis=0)[:, :, classes]
    compute_aos = False
    for anno in dt_annos:                      # :818-823
        if anno['alpha'].shape[0] != 0:
            compute_aos = bool(anno['alpha'][0] != -10)
            break
    diffs = list(difficultys)
    r0 = eval_class(gt_annos, dt_annos, classes, diffs, 0, min_overlaps, compute_aos, overlap_fn)
    ap = dict(bbox=get_mAP(r0["precision"]), aos=get_mAP(r0["orientation"]) if compute_aos else None,
              bev=get_mAP(eval_class(gt_annos, dt_annos, classes, diffs, 1, min_overlaps, False, overlap_fn)["precision"]),
              d3=get_mAP(eval_class(gt_annos, dt_annos, classes, diffs, 2, min_overlaps, False, overlap_fn)["precision"]))
    lines = []
    for j, cls in enumerate(classes):
        for i in range(min_overlaps.shape[0]):
            lines.append("%s AP@%.2f, %.2f, %.2f:" % ((CLASS_TO_NAME[cls],) + tuple(min_overlaps[i, :, j])))
            for key, label in (("bbox", "bbox AP"), ("bev", "bev  AP"), ("d3", "3d   AP")):
                lines.append("%s:%.2f, %.2f, %.2f" % (label, ap[key][j, 0, i], ap[key][j, 1, i], ap[key][j, 2, i]))
            if compute_aos:
                lines.append("aos  AP:%.2f, %.2f, %.2f" % (ap["aos"][j, 0, i], ap["aos"][j, 1, i], ap["aos"][j, 2, i]))
    return "".join(l + "\n" for l in lines), ap


def get_official_eval_result(gt_annos, dt_annos, current_classes, difficultys=[0, 1, 2], overlap_fn=None):
    """Reference signature (kitti_eval.py:791): the printed result table."""
    return official_eval(gt_annos, dt_annos, current_classes, difficultys, overlap_fn)[0]
