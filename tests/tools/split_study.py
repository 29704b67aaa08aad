"""Numerics study: can the dense neck run with TWO fp16 products per algorithmic product instead of
three?  CPU emulation in float64 of what the tensor core would compute (exact products of fp16 pieces, exact sum):

    3-pass (shipped)   x*w ~ xh*wh + xh*wl + xl*wh          hi = fp16(v), lo = fp16((v - hi) * 2048) / 2048
    2-pass A           xh*wh + xl*wh      (weights rounded to fp16, activations split)
    2-pass B           xh*wh + xh*wl      (activations rounded to fp16, weights split)
    2-pass C           like A, weights rounded to bf16-scaled pieces is no better than A by construction; not run
    1-pass             xh*wh

Each layer's INPUT is the exact (float64) output of the previous layer, so the table shows the error one layer adds;
the last rows run the whole neck + heads with the variant in every layer and report the end-to-end error of the
quantities the north-star bounds (class scores, box regressions).  Output: markdown table on stdout.

    python tests/tools/split_study.py [seed]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import torch.nn.functional as F
from oracle import ref_pipeline as O
from sassd_b200.checkpoint import make_synthetic_state_dict
from sassd_b200.synth import synth_cloud

torch.set_num_threads(os.cpu_count() or 8)


def split(v):
    hi = v.to(torch.float32).to(torch.float16).to(torch.float64)
    lo = ((v - hi) * 2048.0).to(torch.float32).to(torch.float16).to(torch.float64) / 2048.0
    return hi, lo


def conv_variant(x, w, pad, variant):
    xh, xl = split(x)
    wh, wl = split(w)
    c = lambda a, b: F.conv2d(a, b, None, padding=pad)
    if variant == "3pass":
        return c(xh, wh) + c(xh, wl) + c(xl, wh)
    if variant == "2passA":
        return c(xh, wh) + c(xl, wh)
    if variant == "2passB":
        return c(xh, wh) + c(xh, wl)
    if variant == "1pass":
        return c(xh, wh)
    return c(x, w)


def bn(x, sd, prefix, eps=1e-3):
    w, b, m, v = (sd[prefix + k].double() for k in (".weight", ".bias", ".running_mean", ".running_var"))
    s = [1, -1, 1, 1]
    return (x - m.view(s)) / torch.sqrt(v.view(s) + eps) * w.view(s) + b.view(s)


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    sd = make_synthetic_state_dict(0, 1)
    v, c, n = O.points_to_voxel(synth_cloud(seed), [0.05, 0.05, 0.1], [0, -40., -3., 70.4, 40., 1.], 5, 20000)
    voxels, coors, num = O.merge_batch([v], [c], [n])
    f, co, shape = O.vxnet_forward(sd, O.simple_voxel(voxels, num), coors, [40, 1600, 1408])
    bev = O.dense_bev(f, co, shape, 1).double()
    variants = ["3pass", "2passA", "2passB", "1pass"]
    print("## per-layer error added by one BEV layer (frame seed %d; exact float64 input, |y| = max |exact output|)\n" % seed)
    print("| layer | |y| | " + " | ".join(variants) + " |")
    print("|---|---:|" + "---:|" * len(variants))
    x = bev
    exact_in = {}
    for i in range(8):
        w = sd["neck.fcn.conv%d.weight" % i].double()
        pad = 1 if i < 7 else 0
        exact_in[i] = x
        y = conv_variant(x, w, pad, "exact")
        errs = [float((conv_variant(x, w, pad, vv) - y).abs().max()) for vv in variants]
        print("| conv%d | %.2f | " % (i, float(y.abs().max())) + " | ".join("%.1e" % e for e in errs) + " |")
        x = torch.relu(bn(y, sd, "neck.fcn.bn%d" % i))
    print("\n## end to end: the variant in all 8 neck layers + the 1x1 heads (float64 elsewhere)\n")
    print("| variant | max err neck output x | max err cls logit | max err sigmoid(cls) | max err box code |")
    print("|---|---:|---:|---:|---:|")

    def run(variant):
        x = bev
        for i in range(8):
            w = sd["neck.fcn.conv%d.weight" % i].double()
            x = torch.relu(bn(conv_variant(x, w, 1 if i < 7 else 0, variant), sd, "neck.fcn.bn%d" % i))
        cls = conv_variant(x, sd["rpn_head.conv_cls.weight"].double(), 0, variant) + sd["rpn_head.conv_cls.bias"].double().view(1, -1, 1, 1)
        box = conv_variant(x, sd["rpn_head.conv_box.weight"].double(), 0, variant) + sd["rpn_head.conv_box.bias"].double().view(1, -1, 1, 1)
        return x, cls, box
    ref = run("exact")
    for vv in variants:
        got = run(vv)
        print("| %s | %.1e | %.1e | %.1e | %.1e |" % (vv, float((got[0] - ref[0]).abs().max()), float((got[1] - ref[1]).abs().max()),
                                                     float((torch.sigmoid(got[1]) - torch.sigmoid(ref[1])).abs().max()),
                                                     float((got[2] - ref[2]).abs().max())))
    print("\n(fp32 itself: an fp32 FFMA chain over K = 2304 terms has ~1e-5 relative error; the bar is 1e-4 absolute on scores and box codes.)")


if __name__ == "__main__":
    main()
