"""Wall time of the KITTI official evaluation on a val-split-sized synthetic set (3769 frames, 3 classes)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_eval import synth_annos  # noqa: E402
from sassd_b200 import kitti_eval as K  # noqa: E402

gts, dts = synth_annos(np.random.default_rng(5), nframes=3769)
K.get_official_eval_result(gts[:50], dts[:50], [0, 1, 2])        # warm-up (context, library)
t0 = time.perf_counter()
ov = [K.calculate_overlaps(gts, dts, m) for m in (0, 1, 2)]
t1 = time.perf_counter()
text = K.get_official_eval_result(gts, dts, [0, 1, 2])
t2 = time.perf_counter()
print("overlaps (3 metrics, %d frames): %.3f s; official result (3 classes x 3 difficulties x 2 overlap sets x 3 metrics): %.2f s"
      % (len(gts), t1 - t0, t2 - t1))
print(text.splitlines()[0], text.splitlines()[3])
