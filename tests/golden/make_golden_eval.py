"""Generate tests/golden/eval.npz from the REFERENCE's KITTI evaluation
(mmdet/core/evaluation/kitti_eval.py + mmdet/core/post_processing/rotate_nms_gpu.py:rotate_iou_gpu_eval) run in
the build container.  The rotated-IoU kernel is numba.cuda code; without a GPU it runs in numba's CUDA simulator:

    NUMBA_ENABLE_CUDASIM=1 python tests/golden/make_golden_eval.py

The fixture holds the synthetic annotations, per-frame overlaps for the three metrics, the AP tables and the
printed official result.
"""
import importlib.util
import os
import sys

import numpy as np
import numpy.ma  # noqa: F401  (before make_golden's np.bool shim)
import numba  # noqa: F401
try:
    import scipy.sparse  # noqa: F401
except Exception:
    pass

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference_mmdet  # noqa: E402

NAMES = ['Car', 'Pedestrian', 'Cyclist', 'Van', 'Person_sitting', 'DontCare']
SIZES = {'Car': (3.9, 1.56, 1.6), 'Van': (5.0, 2.2, 1.9), 'Pedestrian': (0.8, 1.73, 0.6),
         'Person_sitting': (0.8, 1.3, 0.6), 'Cyclist': (1.76, 1.73, 0.6), 'DontCare': (1.0, 1.0, 1.0)}   # l, h, w


def synth_annos(rng, nframes=50):
    gts, dts = [], []
    for _ in range(nframes):
        n = int(rng.integers(3, 10))
        names = rng.choice(NAMES, n, p=[0.45, 0.15, 0.12, 0.1, 0.05, 0.13])
        loc = np.stack([rng.uniform(-20, 20, n), rng.uniform(1.2, 2.0, n), rng.uniform(5, 60, n)], 1)
        dims = np.stack([np.array(SIZES[k]) * rng.uniform(0.85, 1.15, 3) for k in names])
        rot = rng.uniform(-np.pi, np.pi, n)
        # 2-D boxes: height from a pinhole with f = 720, so distant objects fall under the height limits
        h2d = 720.0 * dims[:, 1] / loc[:, 2]
        w2d = 720.0 * np.maximum(dims[:, 0], dims[:, 2]) / loc[:, 2]
        cx = 620 + 720 * loc[:, 0] / loc[:, 2]
        cy = 190 + rng.uniform(-10, 10, n)
        bbox = np.stack([cx - w2d / 2, cy - h2d / 2, cx + w2d / 2, cy + h2d / 2], 1)
        gt = dict(name=np.array(names), truncated=rng.choice([0.0, 0.1, 0.2, 0.4, 0.6], n),
                  occluded=rng.integers(0, 4, n), alpha=rng.uniform(-np.pi, np.pi, n), bbox=bbox,
                  dimensions=dims, location=loc, rotation_y=rot)
        keep = (rng.uniform(size=n) < 0.8) & (names != 'DontCare')
        k = int(keep.sum())
        jit = lambda s, shape: rng.normal(0, s, shape)                        # noqa: E731
        dname = names[keep].copy()
        swap = rng.uniform(size=k) < 0.1
        dname[swap] = rng.choice(['Car', 'Pedestrian', 'Cyclist'], int(swap.sum()))
        nfp = int(rng.integers(0, 4))
        fp_names = rng.choice(['Car', 'Pedestrian', 'Cyclist'], nfp)
        fp_loc = np.stack([rng.uniform(-20, 20, nfp), rng.uniform(1.2, 2.0, nfp), rng.uniform(5, 60, nfp)], 1)
        fp_dims = np.stack([np.array(SIZES[c]) for c in fp_names]) if nfp else np.zeros((0, 3))
        fp_h = 720.0 * fp_dims[:, 1] / fp_loc[:, 2] if nfp else np.zeros((0,))
        fp_cx = 620 + 720 * fp_loc[:, 0] / fp_loc[:, 2] if nfp else np.zeros((0,))
        fp_bbox = np.stack([fp_cx - fp_h, 190 - fp_h / 2, fp_cx + fp_h, 190 + fp_h / 2], 1) if nfp else np.zeros((0, 4))
        dt = dict(name=np.concatenate([dname, fp_names]),
                  truncated=np.zeros(k + nfp), occluded=np.zeros(k + nfp, np.int64),
                  alpha=np.concatenate([gt['alpha'][keep] + jit(0.2, k), rng.uniform(-3, 3, nfp)]),
                  bbox=np.concatenate([bbox[keep] + jit(3.0, (k, 4)), fp_bbox]),
                  dimensions=np.concatenate([dims[keep] * (1 + jit(0.05, (k, 3))), fp_dims]),
                  location=np.concatenate([loc[keep] + jit(0.15, (k, 3)), fp_loc]),
                  rotation_y=np.concatenate([rot[keep] + jit(0.1, k), rng.uniform(-3, 3, nfp)]),
                  score=np.concatenate([rng.uniform(0.3, 1.0, k), rng.uniform(0.05, 0.7, nfp)]))
        gts.append(gt); dts.append(dt)
    return gts, dts


def main():
    assert os.environ.get("NUMBA_ENABLE_CUDASIM") == "1", "run with NUMBA_ENABLE_CUDASIM=1 (no GPU here)"
    import_reference_mmdet()
    for k in list(sys.modules):
        if k.startswith("mmdet.core.post_processing"):
            del sys.modules[k]
    spec = importlib.util.spec_from_file_location(
        "mmdet.core.post_processing.rotate_nms_gpu", "/root/reference/mmdet/core/post_processing/rotate_nms_gpu.py")
    m = importlib.util.module_from_spec(spec); sys.modules[spec.name] = m; spec.loader.exec_module(m)
    spec = importlib.util.spec_from_file_location("ref_kitti_eval", "/root/reference/mmdet/core/evaluation/kitti_eval.py")
    ke = importlib.util.module_from_spec(spec); spec.loader.exec_module(ke)

    rng = np.random.default_rng(11)
    gts, dts = synth_annos(rng)
    out = {"nframes": np.array(len(gts))}
    for i, (g, d) in enumerate(zip(gts, dts)):
        for k, v in g.items():
            out["gt%d_%s" % (i, k)] = v
        for k, v in d.items():
            out["dt%d_%s" % (i, k)] = v
    # raw rotated overlaps on a small box set, all criteria
    b = np.array([[0, 0, 4, 2, 0.3], [1, 0.5, 4, 2, -0.2], [30, 1, 1.8, 0.7, 2.0], [0, 0, 4, 2, 0.3]], np.float32)
    q = np.array([[0.5, 0.2, 3.9, 1.8, 0.1], [10, 10, 1, 1, 0], [30.2, 1.1, 1.7, 0.6, 1.9], [0, 0, 4, 2, 0.3]], np.float32)
    out["probe_boxes"], out["probe_query"] = b, q
    for c in (-1, 0, 1, 2):
        out["probe_crit%d" % c] = m.rotate_iou_gpu_eval(b, q, c)
    for metric in (0, 1, 2):
        ov, _, _, _ = ke.calculate_iou_partly(dts, gts, metric, 50)
        for i, o in enumerate(ov):
            out["ov%d_%d" % (metric, i)] = np.asarray(o)
    classes = [0, 1, 2]
    text = ke.get_official_eval_result(gts, dts, classes)
    out["official_text"] = np.array(text)
    overlap = np.stack([np.array([[0.7, 0.5, 0.5, 0.7, 0.5, 0.7, 0.7, 0.7]] * 3),
                        np.array([[0.7, 0.5, 0.5, 0.7, 0.5, 0.5, 0.5, 0.5], [0.5, 0.25, 0.25, 0.5, 0.25, 0.5, 0.5, 0.5],
                                  [0.5, 0.25, 0.25, 0.5, 0.25, 0.5, 0.5, 0.5]])], 0)[:, :, classes]
    bbox, bev, d3, aos = ke.do_eval_v2(gts, dts, classes, overlap, True, [0, 1, 2])
    out["ap_bbox"], out["ap_bev"], out["ap_d3"], out["ap_aos"] = bbox, bev, d3, aos
    np.savez_compressed(os.path.join(HERE, "eval.npz"), **out)
    print(text)


if __name__ == "__main__":
    main()
