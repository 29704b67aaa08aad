"""``mmdet.models.necks`` mirror for the hot path: SpMiddleFHD = VxNet (sparse 3-D
backbone) + BEVNet (dense BEV convs) — mmdet/models/necks/cmn.py:12-29,102-119,138-282.

Parameter names and shapes equal the reference's so that its
checkpoints load; all arithmetic runs in the sm_100a kernels of csrc/gconv.cu
(BatchNorm folded into each conv's epilogue, eval mode).  The training-only aux head
(cmn.py:44-100,121-135) is out of scope: its Linear parameters exist for checkpoint
compatibility, ``is_test=False`` raises.
"""
import torch
from torch import nn

from . import ops, spconv
from .spconv import fold_bn, _versions


def single_conv(in_channels, out_channels, indice_key=None):
    return spconv.SparseSequential(
        spconv.SubMConv3d(in_channels, out_channels, 1, bias=False, indice_key=indice_key),
        nn.BatchNorm1d(out_channels, eps=1e-3, momentum=0.01),
        nn.ReLU(),
    )


def _subm_block(n, in_channels, out_channels, indice_key):
    layers = []
    for i in range(n):
        layers += [spconv.SubMConv3d(in_channels if i == 0 else out_channels, out_channels, 3, bias=False,
                                     indice_key=indice_key),
                   nn.BatchNorm1d(out_channels, eps=1e-3, momentum=0.01), nn.ReLU()]
    return spconv.SparseSequential(*layers)


def double_conv(in_channels, out_channels, indice_key=None):
    return _subm_block(2, in_channels, out_channels, indice_key)


def triple_conv(in_channels, out_channels, indice_key=None):
    return _subm_block(3, in_channels, out_channels, indice_key)


def stride_conv(in_channels, out_channels, indice_key=None):
    return spconv.SparseSequential(
        spconv.SparseConv3d(in_channels, out_channels, 3, (2, 2, 2), padding=1, bias=False, indice_key=indice_key),
        nn.BatchNorm1d(out_channels, eps=1e-3, momentum=0.01),
        nn.ReLU(),
    )


class VxNet(nn.Module):
    """cmn.py:192-231."""

    def __init__(self, num_input_features):
        super().__init__()
        self.conv0 = double_conv(num_input_features, 16, "subm0")
        self.down0 = stride_conv(16, 32, "down0")
        self.conv1 = double_conv(32, 32, "subm1")
        self.down1 = stride_conv(32, 64, "down1")
        self.conv2 = triple_conv(64, 64, "subm2")
        self.down2 = stride_conv(64, 64, "down2")
        self.conv3 = triple_conv(64, 64, "subm3")
        self.extra_conv = spconv.SparseSequential(
            spconv.SparseConv3d(64, 64, (1, 1, 1), (1, 1, 1), bias=False),
            nn.BatchNorm1d(64, eps=1e-3, momentum=0.01),
            nn.ReLU(),
        )

    def forward(self, x):
        middle = []
        x = self.conv0(x)
        x = self.down0(x)
        x = self.conv1(x)
        middle.append(x)
        x = self.down1(x)
        x = self.conv2(x)
        middle.append(x)
        x = self.down2(x)
        x = self.conv3(x)
        middle.append(x)
        out = self.extra_conv(x)
        return out, middle

    def set_precision(self, precision):
        for m in self.modules():
            if isinstance(m, spconv._SparseConvBase):
                m.precision = precision

    def prebuild_rulebooks(self, x, side_stream, table_stream=None):
        """All seven rulebooks depend on coordinates only, never on features, so they are built beside the feature
        convolutions, which wait on each rulebook's event when they first use it.  Two chains: ``side_stream`` carries
        the coordinate chain - level-0 hash, then per strided conv "mark + single-pass compaction", which also hashes
        the next level's rows (7 kernels end to end); ``table_stream`` fills the seven neighbour tables, each as soon
        as the coordinates and the hash it probes exist.  Round 1 ran all 23 launches back to back on one stream and the
        last table arrived 0.3 ms into a 0.9 ms step - later than the convolutions needed it."""
        main = torch.cuda.current_stream()
        side_stream.wait_stream(main)
        tables = table_stream if table_stream is not None else side_stream
        dev = x.device
        coors, d_rows, shape, cap = x._indices, x.d_rows, x.spatial_shape, x.rows_cap
        with torch.cuda.stream(side_stream):
            index = ops.hash_build(ops.HashIndex(cap, dev), coors, d_rows, x.batch_size, shape, x.status)
        x._index = index
        for lvl in range(4):
            tables.wait_stream(side_stream)              # this level's coordinates and hash are queued
            with torch.cuda.stream(tables):
                nbr, tmask = ops.rulebook_subm(coors, d_rows, shape, index)
                ev = torch.cuda.Event(); ev.record(tables)
                x.indice_dict["subm%d" % lvl] = spconv.Rulebook(nbr, coors, d_rows, shape, index, ev, tmask)
            if lvl == 3:
                break
            D, H, W = ops.conv_out_shape(shape)
            cap = max(1, min(int(cap * x.row_cap_factor), x.batch_size * D * H * W))
            with torch.cuda.stream(side_stream):
                index_out = ops.HashIndex(cap, dev)
                co, dn, so = ops.rulebook_conv_outputs(coors, d_rows, x.batch_size, shape, cap, x.status,
                                                       ws_key="rbconv%d" % lvl, index_out=index_out)
            tables.wait_stream(side_stream)
            with torch.cuda.stream(tables):
                nbr2, tmask2 = ops.rulebook_conv_nbr(co, dn, shape, index)
                ev = torch.cuda.Event(); ev.record(tables)
                x.indice_dict["down%d" % lvl] = spconv.Rulebook(nbr2, co, dn, so, index_out, ev, tmask2)
            coors, d_rows, shape, index = co, dn, so, index_out
        if tables is not side_stream:
            side_stream.wait_stream(tables)              # one join point for the caller
        x._rulebook_stream = side_stream   # keep the stream (and its tensors) alive with the tensor


def pack_conv2d_weight(w, dc_order=None):
    """[Cout, Cin, kh, kw] -> [taps, Cin, Cout] (tap = ky*3+kx).  ``dc_order=(C, D)`` re-orders
    the input channels from the reference's dense() order c*D+d to the internal d*C+c."""
    cout, cin, kh, kw = w.shape
    p = w.detach().permute(2, 3, 1, 0).reshape(kh * kw, cin, cout)
    if dc_order is not None:
        C, D = dc_order
        p = p.reshape(kh * kw, C, D, cout).permute(0, 2, 1, 3).reshape(kh * kw, cin, cout)
    return p.contiguous().float()


def conv2d_nhwc(x, weight_packed, scale, shift, relu, cout, precision=ops.PREC_FP32, out=None, split_out=False):
    """x [B,H,W,Cin] NHWC contiguous -> [B,H,W,cout_stride]; 3x3 (pad 1) when taps == 9, 1x1 when 1.
    A ``ops.SplitMap`` input selects the TMA tensor-core kernel (csrc/conv2d_tma.cu); ``split_out`` then keeps
    the output in split form for the next such layer."""
    if isinstance(x, ops.SplitMap):
        sp, f32 = ops.conv2d_split(x, weight_packed, scale, shift, relu, cout, out_split=split_out,
                                   out_f32=not split_out)
        return sp if split_out else f32
    B, H, W, cin = x.shape
    taps = weight_packed.shape[0]
    stride = (cout + 3) // 4 * 4
    if out is None:
        out = torch.empty((B, H, W, stride), dtype=torch.float32, device=x.device)
        if stride != cout:
            out.zero_()
    ops.gconv(x.view(-1, cin), weight_packed, scale, shift, out.view(-1, out.shape[-1]), mode=ops.GCONV_CONV2D,
              taps=taps, cin=cin, cout=cout, relu=relu, rows_cap=B * H * W, batch=B, H=H, W=W, precision=precision)
    return out


class BEVNet(nn.Module):
    """cmn.py:233-282.  conv{i}/bn{i} hold the reference-named parameters; compute is NHWC."""

    def __init__(self, in_features, num_filters=256):
        super().__init__()
        for i in range(8):
            k = 1 if i == 7 else 3
            setattr(self, "conv%d" % i, nn.Conv2d(in_features if i == 0 else num_filters, num_filters, k,
                                                  padding=k // 2, bias=False))
            setattr(self, "bn%d" % i, nn.BatchNorm2d(num_filters, eps=1e-3, momentum=0.01))
        self.num_filters = num_filters
        self.precision = ops.DEFAULT_PRECISION
        self._packed = {}

    def _weights(self, i, dc_order):
        conv = getattr(self, "conv%d" % i)
        key = (i, dc_order)
        ver = _versions(conv.weight)
        c = self._packed.get(key)
        if c is None or c[0] != ver:
            c = (ver, pack_conv2d_weight(conv.weight, dc_order))
            self._packed[key] = c
        return c[1]

    def forward_nhwc(self, x, dc_order=None):
        """x [B,H,W,Cin] (channel order d*C+c when dc_order=(C,D)).  Returns (x, conv6) NHWC."""
        if self.training:
            raise NotImplementedError("sassd_b200 is inference-only: call .eval()")
        split = isinstance(x, ops.SplitMap)
        for i in range(7):
            scale, shift = fold_bn(getattr(self, "bn%d" % i))
            x = conv2d_nhwc(x, self._weights(i, dc_order if i == 0 else None), scale, shift, True, self.num_filters,
                            self.precision, split_out=split)
        conv6 = x
        scale, shift = fold_bn(self.bn7)
        x = conv2d_nhwc(x, self._weights(7, None), scale, shift, True, self.num_filters, self.precision,
                        split_out=split)
        return x, conv6

    def forward(self, x):
        """Reference signature: x [B, Cin, H, W] -> (x, conv6), both [B, 256, H, W] (channels-last storage)."""
        ops.require_cuda()
        xh = x.permute(0, 2, 3, 1).contiguous()
        y, c6 = self.forward_nhwc(xh)
        return y.permute(0, 3, 1, 2), c6.permute(0, 3, 1, 2)


class SpMiddleFHD(nn.Module):
    """cmn.py:12-29,102-119.  Constructor kwargs as in configs/car_cfg.py:10-15."""

    def __init__(self, output_shape, num_input_features=4, num_hidden_features=128):
        super().__init__()
        self.sparse_shape = list(output_shape)
        self.backbone = VxNet(num_input_features)
        self.fcn = BEVNet(in_features=num_hidden_features, num_filters=256)
        # training-only aux head (cmn.py:27-29): parameters kept so reference checkpoints load completely
        self.point_fc = nn.Linear(160, 64, bias=False)
        self.point_cls = nn.Linear(64, 1, bias=False)
        self.point_reg = nn.Linear(64, 3, bias=False)
        self.row_cap_factor = 4
        self.dense_tma = True             # F16X3: keep the BEV maps as split fp16 planes and feed the convs by TMA
        self.overlap_rulebooks = True     # build the rulebook chain on a side stream, concurrently with the convs
        self._side = None

    def set_precision(self, precision, sparse=None):
        """precision for the dense BEV convs; ``sparse`` (default: same) for the ruled sparse convs."""
        self.backbone.set_precision(precision if sparse is None else sparse)
        self.fcn.precision = precision

    def forward_nhwc(self, voxel_features, coors, batch_size, d_rows=None, status=None):
        """Device-side entry: capacity-sized inputs + row counter; returns NHWC (x, conv6) and the tensor."""
        if self.training:
            raise NotImplementedError("sassd_b200 is inference-only: call .eval()")
        x = spconv.SparseConvTensor(voxel_features, coors, self.sparse_shape, batch_size, d_rows=d_rows, status=status)
        x.row_cap_factor = self.row_cap_factor
        if self.overlap_rulebooks:
            if self._side is None:
                self._side = torch.cuda.Stream(device=x.device)
                self._side2 = torch.cuda.Stream(device=x.device)
            self.backbone.prebuild_rulebooks(x, self._side, self._side2)
        x, middle = self.backbone(x)
        if self.overlap_rulebooks:
            torch.cuda.current_stream().wait_stream(self._side)   # join (also required to end a graph capture)
        C = x._channels
        D, H, W = x.spatial_shape
        if self.fcn.precision == ops.PREC_F16X3 and self.dense_tma:
            # split fp16 planes + TMA-fed tensor-core convs (conv2d_tma.cu); y / conv6 are ops.SplitMap
            if x._split is not None and x._split.shape[2] == C:
                bev = ops.split_rows_to_bev(x._split, x._indices, x.d_rows, C, D, H, W, batch_size)
            else:
                bev = ops.sparse_to_bev_split(x.features_cap(), x._indices, x.d_rows, C, D, H, W, batch_size)
        else:
            feats = x.features_cap()
            bev = torch.zeros((batch_size, H, W, D * C), dtype=torch.float32, device=feats.device)
            ops.sparse_to_bev(feats, x._indices, x.d_rows, C, D, H, W, bev)
        y, conv6 = self.fcn.forward_nhwc(bev, dc_order=(C, D))
        return y, conv6, x

    def forward(self, voxel_features, coors, batch_size, is_test=False, d_rows=None, status=None):
        if not is_test:
            raise NotImplementedError("the auxiliary training branch (cmn.py:121-135) is out of scope")
        y, conv6, x = self.forward_nhwc(voxel_features, coors, batch_size, d_rows, status)
        if isinstance(y, ops.SplitMap):
            y, conv6 = y.float(), conv6.float()
        return y.permute(0, 3, 1, 2), conv6.permute(0, 3, 1, 2)
