"""Pin the oracle (oracle/) against fixtures produced by the REFERENCE's own
Python (tests/golden/make_golden.py, run in the build container)."""
import hashlib
import os

import numpy as np
import torch

from oracle import ref_pipeline as O
from sassd_b200.synth import synth_cloud

VS = [0.05, 0.05, 0.1]
RG = [0, -40., -3., 70.4, 40., 1.]


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode()); h.update(str(a.shape).encode()); h.update(a.tobytes())
    return h.hexdigest()


def _sd(z, prefix):
    return {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}


def test_voxelizer_small_and_truncated(golden_dir):
    z = np.load(os.path.join(golden_dir, "voxelize.npz"))
    for tag in ("small", "small_trunc", "edge"):
        maxv = int(z[tag + "_maxv"]) if tag + "_maxv" in z.files else 20000
        v, c, n = O.points_to_voxel(z[tag + "_points"], VS, RG, 5, maxv)
        assert np.array_equal(c, z[tag + "_coors"]), tag
        assert np.array_equal(n, z[tag + "_num"]), tag
        assert np.array_equal(v, z[tag + "_voxels"]), tag
    assert int(z["small_num"].max()) == 5          # the 5-point cap is exercised
    assert z["small_trunc_coors"].shape[0] == 500  # and so is max_voxels


def test_voxelizer_empty():
    v, c, n = O.points_to_voxel(np.zeros((0, 4), np.float32), VS, RG, 5, 20000)
    assert v.shape == (0, 5, 4) and c.shape == (0, 3) and n.shape == (0,)


def test_voxelizer_full_size_digests(golden_dir):
    z = np.load(os.path.join(golden_dir, "voxelize.npz"))
    for tag, seed, fov in (("full20k", 0, 28.0), ("full45", 1, 45.0)):
        pts = synth_cloud(seed, fov_deg=fov)
        assert pts.shape[0] == int(z[tag + "_npts"])
        assert digest(pts) == str(z[tag + "_points_sha"]), "synthetic cloud generator drifted"
        v, c, n = O.points_to_voxel(pts, VS, RG, 5, 20000)
        assert c.shape[0] == int(z[tag + "_M"])
        assert digest(v, c, n) == str(z[tag + "_sha"])
    assert int(z["full45_M"]) == 20000  # truncation case really truncates


def test_simple_voxel(golden_dir):
    z = np.load(os.path.join(golden_dir, "voxelize.npz"))
    m = np.load(os.path.join(golden_dir, "modules.npz"))
    out = O.simple_voxel(z["small_voxels"], z["small_num"]).numpy()
    assert np.array_equal(out, m["sv_out"])


def test_bevnet(golden_dir):
    m = np.load(os.path.join(golden_dir, "modules.npz"))
    sd = _sd(m, "bev_sd/")
    x, c6 = O.bevnet_forward(sd, torch.from_numpy(m["bev_in"]), prefix="")
    np.testing.assert_allclose(x.numpy(), m["bev_x"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(c6.numpy(), m["bev_conv6"], rtol=0, atol=2e-6)


def test_rpn_head_decode_guided(golden_dir):
    m = np.load(os.path.join(golden_dir, "modules.npz"))
    for ncls in (1, 3):
        p = "head%d_" % ncls
        sd = _sd(m, p + "sd/")
        box, cls, dirp = O.rpn_head_forward(sd, torch.from_numpy(m[p + "x"]), ncls, prefix="")
        np.testing.assert_allclose(box.numpy(), m[p + "box"], atol=1e-6)
        np.testing.assert_allclose(dirp.numpy(), m[p + "dir"], atol=1e-6)
        anchors = torch.from_numpy(m[p + "anchors"])
        dec = O.second_box_decode(torch.from_numpy(m[p + "box"]).view(2, -1, 7), anchors)
        np.testing.assert_allclose(dec.numpy(), m[p + "decoded"], rtol=1e-6, atol=1e-6)
        ga, gl = O.get_guided_anchors(torch.from_numpy(m[p + "box"]), torch.from_numpy(m[p + "cls"]),
                                      torch.from_numpy(m[p + "dir"]), anchors,
                                      torch.from_numpy(m[p + "amask"]), ncls, thr=0.1)
        for b in range(2):
            assert ga[b].shape == m[p + "ga%d" % b].shape
            assert ga[b].shape[0] > 3
            np.testing.assert_allclose(ga[b].numpy(), m[p + "ga%d" % b], rtol=1e-6, atol=1e-6)
            assert np.array_equal(gl[b].numpy(), m[p + "gl%d" % b])


def test_pswarp(golden_dir):
    m = np.load(os.path.join(golden_dir, "modules.npz"))
    sd = _sd(m, "ps_sd/")
    # regenerate the feature map exactly as make_golden.py did (generator state replay)
    from tests.golden_replay import pswarp_feature_and_boxes
    feat, boxes = pswarp_feature_and_boxes()
    for b in range(2):
        np.testing.assert_array_equal(boxes[b].numpy(), m["ps_boxes%d" % b])
    sc = O.pswarp_forward(sd, feat, boxes, (0., 40.), .4, prefix="")
    for b in range(2):
        np.testing.assert_allclose(sc[b].numpy(), m["ps_scores%d" % b], rtol=0, atol=2e-6)
    gx, gy = O.gen_sample_grid(boxes[0][:, [0, 1, 3, 4, 6]])
    np.testing.assert_allclose(gx.numpy(), m["ps_gridx"], atol=1e-5)
    np.testing.assert_allclose(gy.numpy(), m["ps_gridy"], atol=1e-5)


CAR = dict(sizes=[1.6, 3.9, 1.56], anchor_strides=[0.4, 0.4, 1.0], anchor_offsets=[0.2, -39.8, -1.78],
           rotations=[0, 1.57])
PED = dict(CAR, sizes=[0.6, 0.8, 1.73])
CYC = dict(CAR, sizes=[0.6, 1.76, 1.73])


def test_anchors_and_mask(golden_dir):
    a = np.load(os.path.join(golden_dir, "anchors.npz"))
    z = np.load(os.path.join(golden_dir, "voxelize.npz"))
    vs = np.array(VS, np.float32); rg = np.array(RG, np.float32)
    grid = np.round((rg[3:] - rg[:3]) / vs).astype(np.int64)
    for tag, cfgs in (("car", [CAR]), ("multi", [CAR, PED, CYC])):
        anchors, bv = O.make_anchors(cfgs)
        assert anchors.shape[0] == int(a[tag + "_n"])
        assert digest(anchors.astype(np.float32)) == str(a[tag + "_anchors_sha"])
        assert digest(bv.astype(np.float32)) == str(a[tag + "_bv_sha"])
        for ctag in ("small", "edge"):
            mask = O.anchors_mask(z[ctag + "_coors"], bv, vs, rg, grid)
            assert np.array_equal(np.packbits(mask), a["%s_mask_%s" % (tag, ctag)])
    pts = synth_cloud(0)
    _, c, _ = O.points_to_voxel(pts, VS, RG, 5, 20000)
    anchors, bv = O.make_anchors([CAR])
    mask = O.anchors_mask(c, bv, vs, rg, grid)
    assert int(mask.sum()) == int(a["car_mask_full20k_count"])
    assert np.array_equal(np.packbits(mask), a["car_mask_full20k"])
