"""Replays the torch.Generator draws of tests/golden/make_golden.py up to the
PSWarp feature map (4.5 MB, kept out of the fixture) so that it can be rebuilt
bit-identically from the seed."""
import torch


def pswarp_feature_and_boxes():
    g = torch.Generator().manual_seed(0)

    def bn_draws(channels_list):
        for c in channels_list:
            torch.randn(c, generator=g); torch.rand(c, generator=g)
            torch.rand(c, generator=g); torch.randn(c, generator=g)
    bn_draws([16] * 8)                       # BEVNet(20,16): bn0..bn7
    torch.randn(2, 20, 12, 10, generator=g)  # BEVNet input
    for ncls in (1, 3):                      # SSDRotateHead fixtures
        torch.randn(2, 16, 6, 5, generator=g)
        na = ncls * 6 * 5 * 2
        torch.randn(2, na, 7, generator=g)
        torch.rand(2, na, generator=g)
    bn_draws([28])                           # PSWarpHead.convs[1]
    feat = torch.randn(2, 16, 200, 176, generator=g)
    boxes = []
    for b in range(2):
        k = 37 + 5 * b
        bx = torch.zeros(k, 7)
        bx[:, 0] = torch.rand(k, generator=g) * 76 - 3
        bx[:, 1] = torch.rand(k, generator=g) * 86 - 43
        bx[:, 2] = -1.0
        bx[:, 3] = 1.6 + torch.randn(k, generator=g) * 0.1
        bx[:, 4] = 3.9 + torch.randn(k, generator=g) * 0.3
        bx[:, 5] = 1.56
        bx[:, 6] = (torch.rand(k, generator=g) - 0.5) * 8
        boxes.append(bx)
    return feat, boxes
