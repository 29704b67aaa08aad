"""Generate tests/golden/results.npz from the REFERENCE's result formatter
(mmdet/core/bbox/transforms.py:225 kitti_bbox2results, kitti_utils.Calibration)
run in the build container.  /root/reference is absent on the GPU box, so the
fixture is committed.

    python tests/golden/make_golden_results.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference_mmdet  # noqa: E402

CALIB_TXT = """P0: 7.215377e+02 0.0 6.095593e+02 0.0 0.0 7.215377e+02 1.728540e+02 0.0 0.0 0.0 1.0 0.0
P1: 7.215377e+02 0.0 6.095593e+02 -3.875744e+02 0.0 7.215377e+02 1.728540e+02 0.0 0.0 0.0 1.0 0.0
P2: 7.215377e+02 0.0 6.095593e+02 4.485728e+01 0.0 7.215377e+02 1.728540e+02 2.163791e-01 0.0 0.0 1.0 2.745884e-03
P3: 7.215377e+02 0.0 6.095593e+02 -3.395242e+02 0.0 7.215377e+02 1.728540e+02 2.199936e+00 0.0 0.0 1.0 2.729905e-03
R0_rect: 9.999239e-01 9.837760e-03 -7.445048e-03 -9.869795e-03 9.999421e-01 -4.278459e-03 7.402527e-03 4.351614e-03 9.999631e-01
Tr_velo_to_cam: 7.533745e-03 -9.999714e-01 -6.166020e-04 -4.069766e-03 1.480249e-02 7.280733e-04 -9.998902e-01 -7.631618e-02 9.998621e-01 7.523790e-03 1.480755e-02 -2.717806e-01
Tr_imu_to_velo: 9.999976e-01 7.553071e-04 -2.035826e-03 -8.086759e-01 -7.854027e-04 9.998898e-01 -1.482298e-02 3.195559e-01 2.024406e-03 1.482454e-02 9.998881e-01 -7.997231e-01
"""


def make_cases(rng):
    cases = []
    for n in (0, 1, 7, 40):
        boxes = np.zeros((n, 7), np.float32)
        boxes[:, 0] = rng.uniform(-5, 70, n)       # some behind / off image
        boxes[:, 1] = rng.uniform(-40, 40, n)
        boxes[:, 2] = rng.uniform(-2.5, 0.0, n)
        boxes[:, 3] = rng.uniform(1.4, 2.0, n)
        boxes[:, 4] = rng.uniform(3.0, 5.0, n)
        boxes[:, 5] = rng.uniform(1.3, 1.9, n)
        boxes[:, 6] = rng.uniform(-8, 8, n)
        scores = rng.uniform(0.3, 1.0, n).astype(np.float32)
        labels = rng.integers(0, 3, n)
        cases.append((boxes, scores, labels))
    return cases


def main():
    import_reference_mmdet()
    from mmdet.core.bbox.transforms import kitti_bbox2results
    from mmdet.datasets.kitti_utils import Calibration
    with tempfile.NamedTemporaryFile('w', suffix='.txt', delete=False) as fh:
        fh.write(CALIB_TXT)
        path = fh.name
    calib = Calibration(path)
    os.unlink(path)
    names = ['Car', 'Pedestrian', 'Cyclist']
    out = {'calib_txt': np.array(CALIB_TXT), 'ncases': np.array(0)}
    rng = np.random.default_rng(7)
    for ci, (boxes, scores, labels) in enumerate(make_cases(rng)):
        meta = dict(calib=calib, sample_idx=100 + ci, img_shape=(375, 1242, 3))
        out['c%d_boxes' % ci] = boxes.copy()
        out['c%d_scores' % ci] = scores
        out['c%d_labels' % ci] = labels
        res = kitti_bbox2results(boxes.copy(), scores, labels, meta, names)
        for k, v in res.items():
            out['c%d_out_%s' % (ci, k)] = np.asarray(v)
        if len(res['name']):
            import tools.kitti_common as kitti
            out['c%d_lines' % ci] = np.array(kitti.annos_to_kitti_label(res))
            out['c%d_line0_scored' % ci] = np.array(kitti.kitti_result_line(
                {k: res[k][0] for k in ('name', 'truncated', 'occluded', 'alpha', 'bbox', 'dimensions', 'location',
                                        'rotation_y', 'score')}, precision=2))
        print(ci, {k: np.asarray(v).shape for k, v in res.items()})
    out['ncases'] = np.array(ci + 1)
    np.savez_compressed(os.path.join(HERE, 'results.npz'), **out)


if __name__ == '__main__':
    main()
