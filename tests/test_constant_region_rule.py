"""The rule behind the constant-region tile skipping of the dense neck, checked on the CPU with
an fp64 torch model: for a BEV map scattered from a sparse tensor, a conv layer whose output is `reach` 3x3
convolutions away from that map is constant on every 8x16 tile whose distance to the nearest active cell exceeds
`reach` - except, from the second conv on, the tiles on the image border (zero padding != constant).  The numpy
functions below restate csrc/common.cuh:sassd_mark_conv2d_tiles and csrc/conv2d_tma.cu:tile_is_constant."""
import numpy as np
import torch

TH, TW, DMAX, FAR = 8, 16, 9, 1 << 20


def tile_distances(active_yx, H, W):
    ty, tx = (H + TH - 1) // TH, (W + TW - 1) // TW
    dist = np.full((ty, tx), FAR, np.int64)
    for y, x in active_yx:
        for j in range(max(y - DMAX, 0) // TH, min(y + DMAX, H - 1) // TH + 1):
            for i in range(max(x - DMAX, 0) // TW, min(x + DMAX, W - 1) // TW + 1):
                y0, x0 = j * TH, i * TW
                dy = max(y0 - y, y - (y0 + TH - 1), 0)
                dx = max(x0 - x, x - (x0 + TW - 1), 0)
                dist[j, i] = min(dist[j, i], max(dy, dx))
    return dist


def tile_is_constant(dist, reach):
    const = dist > reach
    if reach >= 2:
        const[0, :] = const[-1, :] = False
        const[:, 0] = const[:, -1] = False
    return const


def skipping_valid(H, W, reach):
    """Restates sassd_b200.ops.tile_skipping_valid / the guard of sassd_conv2d_f16x3_occ."""
    last_h, last_w = H - (H - 1) // TH * TH, W - (W - 1) // TW * TW
    return reach <= DMAX and min(last_h, last_w, TH, TW) >= reach - 1


def test_guard_matches_the_product():
    from sassd_b200 import ops
    for H, W in ((200, 176), (72, 112), (70, 100), (9, 17), (64, 33)):
        for reach in range(0, 12):
            assert ops.tile_skipping_valid(H, W, reach) == skipping_valid(H, W, reach)


def test_constant_tiles_rule_on_a_conv_chain():
    _chain(72, 112)


def test_constant_tiles_rule_with_partial_edge_tiles():
    """H, W not multiples of the tile: the rule holds exactly as long as the guard allows skipping, and the guard is
    not vacuous - past it a skipped tile really is non-constant (thin edge tile, padding influence spills inward)."""
    broke = _chain(70, 100, expect_break=True)
    assert broke is not None and not skipping_valid(70, 100, broke)


def _chain(H, W, expect_break=False):
    torch.manual_seed(0)
    rs = np.random.RandomState(0)
    C = 6
    active = {(int(rs.randint(0, 20)), int(rs.randint(0, 30))) for _ in range(25)} | {(H - 1, W - 1), (40, 60)}
    x = torch.zeros(1, C, H, W, dtype=torch.float64)
    for y, xx in active:
        x[0, :, y, xx] = torch.randn(C, dtype=torch.float64)
    dist = tile_distances(sorted(active), H, W)
    layers = [(3, C, 8), (3, 8, 8), (3, 8, 8), (1, 8, 8), (3, 8, 5), (3, 5, 5), (3, 5, 5), (3, 5, 5), (3, 5, 5)]
    reach, const_in = 0, torch.zeros(C, dtype=torch.float64)
    skipped = 0
    for k, ci, co in layers:
        w = torch.randn(co, ci, k, k, dtype=torch.float64) * 0.4
        b = torch.randn(co, dtype=torch.float64) * 0.5 + 0.3
        x = torch.relu(torch.nn.functional.conv2d(x, w, b, padding=k // 2))
        # the layer's constant: the same conv on an all-constant map, read at an interior pixel
        rep = const_in.view(1, ci, 1, 1).expand(1, ci, 3 * TH, 3 * TW)
        const_out = torch.relu(torch.nn.functional.conv2d(rep, w, b, padding=k // 2))[0, :, 3 * TH // 2, 3 * TW // 2]
        reach += 1 if k == 3 else 0
        valid = skipping_valid(H, W, reach)
        assert valid or expect_break
        cst = tile_is_constant(dist.copy(), reach)
        for j in range(cst.shape[0]):
            for i in range(cst.shape[1]):
                if cst[j, i]:
                    patch = x[0, :, j * TH:(j + 1) * TH, i * TW:(i + 1) * TW]
                    same = torch.allclose(patch, const_out.view(co, 1, 1).expand_as(patch), rtol=0, atol=1e-12)
                    if not same and not valid:
                        return reach          # outside the guard the rule may (and here does) fail
                    assert same, "layer reach %d tile (%d,%d) is not constant" % (reach, j, i)
                    skipped += 1
        const_in = const_out
    assert skipped > 50          # the rule actually skips work on this map
    return None


def test_rule_is_tight_at_the_boundary():
    """A tile exactly `reach` pixels from an active cell is NOT constant (its nearest pixel sees the cell)."""
    H, W = 48, 64
    for reach in (1, 2, 3):
        y, x = 8 + 7 + reach, 20          # `reach` rows below tile row 1 (rows 8..15)
        dist = tile_distances([(y, x)], H, W)
        assert dist[1, 1] == reach and not tile_is_constant(dist.copy(), reach)[1, 1]
        xin = torch.zeros(1, 1, H, W, dtype=torch.float64)
        xin[0, 0, y, x] = 1.0
        out = xin
        for _ in range(reach):
            out = torch.nn.functional.conv2d(out, torch.ones(1, 1, 3, 3, dtype=torch.float64), padding=1)
        assert out[0, 0, 15, 20] != 0          # the influence reaches the tile's last row
        assert (out[0, 0, 8:15, 16:32] == 0).all() or reach > 1
