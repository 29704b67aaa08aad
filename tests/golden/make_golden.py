"""Generate tests/golden/*.npz by running the REFERENCE's own Python
(/root/reference, read-only) in the build container.  Run once; the fixtures are
committed because /root/reference does not exist on the GPU box.

    python tests/golden/make_golden.py

What runs from the reference (SURVEY.md §8c / Appendix B recipe):
  * mmdet/ops/points_op/points_ops.py  (numba voxelizer; loaded by file path)
  * mmdet/models/backbones/vxnet.py:SimpleVoxel
  * mmdet/models/necks/cmn.py:BEVNet
  * mmdet/models/single_stage_heads/ssd_rotate_head.py: SSDRotateHead (+decode,
    guided anchors), PSWarpHead (+gen_sample_grid, grid_sample wrapper)
  * mmdet/core/anchor/anchor3d_generator.py:AnchorGeneratorStride
  * mmdet/core/bbox3d/geometry.py: rbbox2d_to_near_bbox,
    sparse_sum_for_anchors_mask, fused_get_anchors_area
spconv (third-party, absent) and iou3d_cuda (CUDA-only) cannot run here.
"""
import collections
import collections.abc
import hashlib
import importlib.util
import os
import sys
import unittest.mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode()); h.update(str(a.shape).encode()); h.update(a.tobytes())
    return h.hexdigest()


def load_ref_points_ops():
    spec = importlib.util.spec_from_file_location(
        "ref_points_ops", os.path.join(REF, "mmdet/ops/points_op/points_ops.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def import_reference_mmdet():
    collections.Sequence = collections.abc.Sequence
    np.bool = bool
    _mg = np.meshgrid
    np.meshgrid = lambda *a, **k: list(_mg(*a, **k))
    sys.path.insert(0, REF)
    for name in ["mmcv", "mmcv.runner", "mmcv.runner.log_buffer", "mmcv.parallel", "mmcv.cnn", "spconv",
                 "pycocotools", "pycocotools.coco", "pycocotools.cocoeval", "pycocotools.mask", "terminaltables",
                 "shapely", "shapely.geometry", "mayavi", "mayavi.mlab", "matplotlib", "matplotlib.pyplot",
                 "skimage", "skimage.io", "imageio", "fire",
                 "mmdet.ops.iou3d.iou3d_cuda", "mmdet.ops.pointnet2.pointnet2_cuda",
                 "mmdet.ops.points_op.points_op_cpu", "mmdet.core.post_processing.rotate_nms_gpu"]:
        m = unittest.mock.MagicMock(name=name)
        m.__path__ = []
        m.__spec__ = None
        sys.modules[name] = m

    class _Base:
        def __init__(self, *a, **k):
            pass
    sys.modules["mmcv.runner"].Hook = _Base
    sys.modules["mmcv.runner"].OptimizerHook = _Base


def randomize_bn(module, gen):
    for m in module.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=gen) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=gen) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=gen) + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=gen) * 0.1)


def sd_np(module):
    return {k: v.detach().cpu().numpy() for k, v in module.state_dict().items() if "num_batches" not in k}


def main():
    from sassd_b200.synth import synth_cloud
    VS = [0.05, 0.05, 0.1]
    RG = [0, -40., -3., 70.4, 40., 1.]

    # ---------------- voxelizer ----------------
    ops = load_ref_points_ops()
    cloud0 = synth_cloud(0)
    out = {}
    rs = np.random.RandomState(0)
    sub = cloud0[rs.permutation(cloud0.shape[0])[:3000]].copy()
    # duplicate some points / crowd a few voxels so that the 5-point cap is hit
    crowd = sub[:40].repeat(8, axis=0) + rs.uniform(-0.004, 0.004, (320, 4)).astype(np.float32)
    small = np.concatenate([sub, crowd.astype(np.float32)], 0)
    small = small[rs.permutation(small.shape[0])]
    for tag, pts, maxv in [("small", small, 20000), ("small_trunc", small, 500)]:
        v, c, n = ops.points_to_voxel(pts, VS, RG, 5, True, maxv)
        out[tag + "_points"] = pts
        out[tag + "_voxels"] = v; out[tag + "_coors"] = c; out[tag + "_num"] = n
        out[tag + "_maxv"] = np.int32(maxv)
    # edge cases: empty cloud, everything out of range, boundary-hugging coordinates
    edge = np.array([[0.0, -40.0, -3.0, .5], [70.4, 0, 0, .5], [70.39999, 39.99999, 0.99999, .1],
                     [-0.00001, 0, 0, .2], [0.05, -39.95, -2.9, .3], [0.049999997, -39.95, -2.9, .4],
                     [35.2, 0.0, -1.0, .6], [35.2, 0.0, -1.0, .7], [1e9, 0, 0, 0], [10, -50, 0, 0]], np.float32)
    grid_pts = (np.arange(0, 400, dtype=np.float32)[:, None] * np.array([0.05, 0.05, 0.01, 0], np.float32)
                + np.array([3.0, -2.0, -3.0, 0.5], np.float32)).astype(np.float32)
    edge = np.concatenate([edge, grid_pts], 0)
    v, c, n = ops.points_to_voxel(edge, VS, RG, 5, True, 20000)
    out.update(edge_points=edge, edge_voxels=v, edge_coors=c, edge_num=n)
    v, c, n = ops.points_to_voxel(np.zeros((0, 4), np.float32), VS, RG, 5, True, 20000)
    out.update(empty_M=np.int32(c.shape[0]))
    # full-size clouds: digests only
    for tag, fov, maxv in [("full20k", 28.0, 20000), ("full45", 45.0, 20000)]:
        pts = synth_cloud(1 if tag == "full45" else 0, fov_deg=fov)
        v, c, n = ops.points_to_voxel(pts, VS, RG, 5, True, maxv)
        out[tag + "_npts"] = np.int64(pts.shape[0])
        out[tag + "_M"] = np.int64(c.shape[0])
        out[tag + "_points_sha"] = np.array(digest(pts))
        out[tag + "_sha"] = np.array(digest(v, c, n))
    np.savez_compressed(os.path.join(HERE, "voxelize.npz"), **out)
    print("voxelize:", {k: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in out.items()
                        if "sha" not in k and "points" not in k and "voxels" not in k})

    # ---------------- modules ----------------
    import_reference_mmdet()
    from mmdet.models.backbones.vxnet import SimpleVoxel
    from mmdet.models.necks.cmn import BEVNet
    from mmdet.models.single_stage_heads import ssd_rotate_head as H
    from mmdet.core.anchor.anchor3d_generator import AnchorGeneratorStride
    from mmdet.core.bbox3d import geometry as G

    g = torch.Generator().manual_seed(0)
    out = {}
    with torch.no_grad():
        # SimpleVoxel on the small voxel set
        vz = np.load(os.path.join(HERE, "voxelize.npz"))
        sv = SimpleVoxel(num_input_features=4)
        out["sv_out"] = sv(torch.from_numpy(vz["small_voxels"]), torch.from_numpy(vz["small_num"])).numpy()

        # BEVNet, reduced width
        torch.manual_seed(1)
        net = BEVNet(in_features=20, num_filters=16).eval()
        randomize_bn(net, g)
        x = torch.randn(2, 20, 12, 10, generator=g)
        y, c6 = net(x)
        out["bev_in"] = x.numpy(); out["bev_x"] = y.numpy(); out["bev_conv6"] = c6.numpy()
        for k, v in sd_np(net).items():
            out["bev_sd/" + k] = v

        # SSDRotateHead, num_class 1 and 3
        for ncls in (1, 3):
            torch.manual_seed(2 + ncls)
            head = H.SSDRotateHead(num_class=ncls, num_output_filters=16, num_anchor_per_loc=2,
                                   use_sigmoid_cls=True, encode_rad_error_by_sin=True,
                                   use_direction_classifier=True, box_code_size=7).eval()
            Hh, Ww = 6, 5
            x = torch.randn(2, 16, Hh, Ww, generator=g)
            box, cls, dirp = head(x)
            na = ncls * Hh * Ww * 2
            anchors = torch.randn(2, na, 7, generator=g)
            anchors[..., 3:6] = anchors[..., 3:6].abs() + 0.5
            amask = torch.rand(2, na, generator=g) > 0.3
            # spread the class logits so that the 0.1 threshold separates anchors
            cls = cls * 6.0 - 1.0
            ga, gl = head.get_guided_anchors(box.clone(), cls.clone(), dirp.clone(), anchors, amask, None, None, thr=.1)
            p = "head%d_" % ncls
            out[p + "x"] = x.numpy(); out[p + "box"] = box.numpy(); out[p + "cls"] = cls.numpy()
            out[p + "dir"] = dirp.numpy(); out[p + "anchors"] = anchors.numpy(); out[p + "amask"] = amask.numpy()
            out[p + "decoded"] = H.second_box_decode(box.view(2, -1, 7), anchors).numpy()
            for b in range(2):
                out[p + "ga%d" % b] = ga[b].numpy(); out[p + "gl%d" % b] = gl[b].numpy()
            for k, v in sd_np(head).items():
                out[p + "sd/" + k] = v

        # PSWarpHead
        torch.manual_seed(7)
        ps = H.PSWarpHead(grid_offsets=(0., 40.), featmap_stride=.4, in_channels=16, num_class=1, num_parts=28).eval()
        randomize_bn(ps, g)
        feat = torch.randn(2, 16, 200, 176, generator=g)
        boxes = []
        for b in range(2):
            k = 37 + 5 * b
            bx = torch.zeros(k, 7)
            bx[:, 0] = torch.rand(k, generator=g) * 76 - 3      # some partly outside the map
            bx[:, 1] = torch.rand(k, generator=g) * 86 - 43
            bx[:, 2] = -1.0
            bx[:, 3] = 1.6 + torch.randn(k, generator=g) * 0.1
            bx[:, 4] = 3.9 + torch.randn(k, generator=g) * 0.3
            bx[:, 5] = 1.56
            bx[:, 6] = (torch.rand(k, generator=g) - 0.5) * 8
            boxes.append(bx)
        sc = ps(feat, boxes, is_test=True)
        out["ps_feat_seed"] = np.int64(7)
        out["ps_feat"] = feat[:, :, ::1].numpy().astype(np.float32)
        for b in range(2):
            out["ps_boxes%d" % b] = boxes[b].numpy(); out["ps_scores%d" % b] = sc[b].numpy()
        gx, gy = H.gen_sample_grid(boxes[0][:, [0, 1, 3, 4, 6]].clone(), grid_offsets=(0., 40.), spatial_scale=2.5)
        out["ps_gridx"] = gx.numpy(); out["ps_gridy"] = gy.numpy()
        for k, v in sd_np(ps).items():
            out["ps_sd/" + k] = v
    # the PSWarp feature map is 2*16*200*176*4 = 4.5 MB: keep it out of the fixture, regenerate from seed
    del out["ps_feat"]
    np.savez_compressed(os.path.join(HERE, "modules.npz"), **out)
    print("modules:", len(out), "arrays")

    # ---------------- anchors + anchors mask ----------------
    out = {}
    car = dict(sizes=[1.6, 3.9, 1.56], anchor_strides=[0.4, 0.4, 1.0], anchor_offsets=[0.2, -39.8, -1.78],
               rotations=[0, 1.57])
    ped = dict(car, sizes=[0.6, 0.8, 1.73])
    cyc = dict(car, sizes=[0.6, 1.76, 1.73])
    fms = [1, 200, 176]
    for tag, cfgs in [("car", [car]), ("multi", [car, ped, cyc])]:
        anchors = np.concatenate([AnchorGeneratorStride(**c)(fms).reshape(-1, 7) for c in cfgs], 0)
        bv = G.rbbox2d_to_near_bbox(anchors[..., [0, 1, 3, 4, 6]])
        out[tag + "_anchors_sha"] = np.array(digest(anchors.astype(np.float32)))
        out[tag + "_bv_sha"] = np.array(digest(bv.astype(np.float32)))
        out[tag + "_anchors_head"] = anchors[:6]; out[tag + "_anchors_tail"] = anchors[-3:]
        out[tag + "_n"] = np.int64(anchors.shape[0])
        for ctag in ("small", "edge"):
            coors = vz[ctag + "_coors"]
            vs = np.array(VS, np.float32); rg = np.array(RG, np.float32)
            grid = np.round((rg[3:] - rg[:3]) / vs).astype(np.int64)
            dm = G.sparse_sum_for_anchors_mask(coors, tuple(grid[::-1][1:]))
            dm = dm.cumsum(0); dm = dm.cumsum(1)
            area = G.fused_get_anchors_area(dm, bv, vs, rg, grid)
            out["%s_mask_%s" % (tag, ctag)] = np.packbits(area > 1)
            out["%s_area_%s_sha" % (tag, ctag)] = np.array(digest(area.astype(np.float32)))
    # full-size frame mask (car), digest only
    pts = synth_cloud(0)
    v, c, n = ops.points_to_voxel(pts, VS, RG, 5, True, 20000)
    anchors = AnchorGeneratorStride(**car)(fms).reshape(-1, 7)
    bv = G.rbbox2d_to_near_bbox(anchors[..., [0, 1, 3, 4, 6]])
    vs = np.array(VS, np.float32); rg = np.array(RG, np.float32)
    grid = np.round((rg[3:] - rg[:3]) / vs).astype(np.int64)
    dm = G.sparse_sum_for_anchors_mask(c, tuple(grid[::-1][1:])).cumsum(0).cumsum(1)
    m = G.fused_get_anchors_area(dm, bv, vs, rg, grid) > 1
    out["car_mask_full20k"] = np.packbits(m)
    out["car_mask_full20k_count"] = np.int64(m.sum())
    np.savez_compressed(os.path.join(HERE, "anchors.npz"), **out)
    print("anchors: car n=%d multi n=%d, full mask count %d" % (out["car_n"], out["multi_n"], m.sum()))


if __name__ == "__main__":
    main()
