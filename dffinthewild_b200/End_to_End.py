"""Drop-in `End_to_End.Network` (reference End_to_End/End_to_End.py:8-17): the alignment network `FlowNetwork(8)` followed by
`DFF_net`, both on hand-written sm_100a CUDA through the C-ABI library.

Interface contract (reference `test_real_scenes.py:11,20-34`): `Network()` takes no arguments, has the children `DFF_net` and
`optical_flow_aggregation`; `forward(FS, focus_dists, FOVs)` returns `(mid_out, pred1, pred2, pred3, warped_FS)`; the
`state_dict` has the reference's 522 keys.  `torch.manual_seed(k); Network()` reproduces the reference's as-built weights
(`DFF_net` with its N(0, sqrt(2/n)) re-draw first, then `FlowNetwork` with PyTorch's default initialisation).

`FlowNetwork.forward` (reference :63-104) runs in eval mode only in this build (the reference ships no training script for it)
and always in fp32 — its output is a sub-pixel warp: six `resnet_block_2d_OF` blocks and three alignment heads (`dff_conv3d`,
BatchNorm folded into the epilogue), `dff_fov_warp_cl` on the feature volumes, `dff_pair_volume` (last-slice ‖ slice ‖ flow input
of a head), `dff_spatial_mean_accum` (AdaptiveAvgPool3d((S,1,1)), the 0.001 factor and the running sum of alpha), and finally
`dff_fov_warp` on the focal stack itself.  Like the reference, it requires S == 10 and is bug-compatible for B > 1
(sample 0's scale correction is applied to every sample).
"""
import ctypes

import torch
import torch.nn as nn

from . import runtime as _rt
from .Depth_Estimation_Network import DFF_net as _DFFNetBase
from .Depth_Estimation_Network import _Container, _conv_bn, _relu

_P = ctypes.c_void_p


class _ResBlock2dOF(_Container):
    """`resnet_block_2d_OF` (reference :135-145): relu(feature(x) + convbn(relu(convbn(x))))."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv = nn.Sequential(
            _conv_bn(cin, cout, (1, 3, 3), (1, stride, stride), (0, 1, 1), (1, 1, 1)),
            _relu(),
            _conv_bn(cout, cout, (1, 3, 3), (1, 1, 1), (0, 1, 1), (1, 1, 1)),
        )
        self.feature = nn.Conv3d(cin, cout, 1, (1, stride, stride), 0, bias=False)
        self.relu = _relu()


def _head(cin, c):
    """One alignment head (reference :31-61): 3 x (conv 1x3x3 + BN + ReLU), conv 1x3x3 -> 3 with bias, AdaptiveAvgPool3d((10,1,1))."""
    return nn.Sequential(
        _conv_bn(cin, c, (1, 3, 3), 1, (0, 1, 1)), _relu(),
        _conv_bn(c, c, (1, 3, 3), 1, (0, 1, 1)), _relu(),
        _conv_bn(c, c, (1, 3, 3), 1, (0, 1, 1)), _relu(),
        nn.Conv3d(c, 3, (1, 3, 3), stride=1, padding=(0, 1, 1)),
        nn.AdaptiveAvgPool3d((10, 1, 1)),
    )


class FlowNetwork(nn.Module):
    """Alignment network (reference `FlowNetwork`, :18-104): parameter containers + a forward made of C-ABI calls."""

    def __init__(self, inplanes):
        super().__init__()
        p = inplanes
        self.OF_feature = nn.Sequential(_ResBlock2dOF(3, p, 1), _ResBlock2dOF(p, p, 1))
        self.OF_feature1 = nn.Sequential(_ResBlock2dOF(p, 2 * p, 2), _ResBlock2dOF(2 * p, 2 * p, 1))
        self.OF_feature2 = nn.Sequential(_ResBlock2dOF(2 * p, 4 * p, 2), _ResBlock2dOF(4 * p, 4 * p, 1))
        self.conv1 = _head(8 * p + 2, 8 * p)
        self.conv2 = _head(4 * p + 2, 4 * p)
        self.conv3 = _head(2 * p + 2, 2 * p)

    def forward(self, FS, FOVs):
        return flow_forward(self, FS, FOVs)


class DFF_net(_DFFNetBase):
    """The End-to-End copy of `DFF_net` (reference :147-259) returns its (aligned) input as a fifth output."""

    def forward(self, FS, focus_dists):
        return tuple(_rt.dff_net_forward(self, FS, focus_dists)) + (FS,)


class Network(nn.Module):
    def __init__(self):
        super().__init__()
        self.DFF_net = DFF_net()
        self.optical_flow_aggregation = FlowNetwork(8)

    def forward(self, FS, focus_dists, FOVs):
        FS = self.optical_flow_aggregation(FS, FOVs)
        return self.DFF_net(FS, focus_dists)


# ---------------------------------------------------------------------------------------------------------------
# FlowNetwork.forward on the C-ABI (fp32, channels-last volumes)
# ---------------------------------------------------------------------------------------------------------------
def _p(t):
    return _P(t.data_ptr()) if t is not None else _P(0)


def _st(dev):
    return _P(torch.cuda.current_stream(dev).cuda_stream)


def _fold(bn):
    """BatchNorm3d (eval) -> per-channel scale / shift of the conv epilogue."""
    scale = (bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps))
    shift = bn.bias.detach().double() - bn.running_mean.detach().double() * scale
    return scale.float().contiguous(), shift.float().contiguous()


def _conv_cl(x, weight, stride=1, scale=None, shift=None, res_pre=None, relu=False, cin_pad_to=None):
    """One fused conv on a channels-last fp32 volume (B,S,H,W,C) through `dff_conv3d` (FFMA parity path)."""
    l = _rt.lib()
    dev = x.device
    B, S, IH, IW, C0 = x.shape
    w = weight.detach().float()
    if cin_pad_to is not None and w.shape[1] < cin_pad_to:   # stored channels beyond the layer's Cin carry zero weights
        w = torch.cat([w, w.new_zeros(w.shape[0], cin_pad_to - w.shape[1], *w.shape[2:])], 1)
    w = w.contiguous()
    Cout = w.shape[0]
    kd, kh, kw = w.shape[2:]
    out = torch.empty((B, S, IH // stride, IW // stride, Cout), dtype=torch.float32, device=dev)
    scratch = torch.empty(l.dff_conv3d_scratch_bytes(C0, Cout, kd, kh, kw), dtype=torch.uint8, device=dev)
    _rt.check(l.dff_conv3d(_p(x), C0, None, 0, B, S, IH, IW, _p(w), Cout, kd, kh, kw, stride, 1, 0, _p(scale), _p(shift), _p(res_pre),
                           None, 1 if relu else 0, _p(out), _rt.FP32, 0, _p(scratch), dev.index, _st(dev)))
    return out


def _block(m, x, cin_pad_to=None):
    stride = m.feature.stride[1]
    s0, b0 = _fold(m.conv[0][1])
    s1, b1 = _fold(m.conv[2][1])
    t = _conv_cl(x, m.conv[0][0].weight, stride, s0, b0, relu=True, cin_pad_to=cin_pad_to)
    f = _conv_cl(x, m.feature.weight, stride, cin_pad_to=cin_pad_to)
    return _conv_cl(t, m.conv[2][0].weight, 1, s1, b1, res_pre=f, relu=True)


def _warp_cl(x, alpha, fov):
    l = _rt.lib()
    B, S, H, W, C = x.shape
    out = torch.empty_like(x)
    _rt.check(l.dff_fov_warp_cl(_p(x), _p(alpha), _p(fov), B, C, S, H, W, _p(out), _rt.FP32, x.device.index, _st(x.device)))
    return out


def _align_head(head, feat, alpha, fov):
    """warp -> pair volume -> 3 x conv+BN+ReLU -> conv(+bias) -> per-slice spatial mean, scaled and added to alpha."""
    l = _rt.lib()
    dev = feat.device
    warped = _warp_cl(feat, alpha, fov)
    B, S, H, W, C = warped.shape
    vol = torch.empty((B, S, H, W, 2 * C + 8), dtype=torch.float32, device=dev)
    _rt.check(l.dff_pair_volume(_p(warped), _p(alpha), _p(fov), B, C, S, H, W, _p(vol), _rt.FP32, dev.index, _st(dev)))
    t = vol
    for i in (0, 2, 4):
        sc, sh = _fold(head[i][1])
        t = _conv_cl(t, head[i][0].weight, 1, sc, sh, relu=True, cin_pad_to=t.shape[-1])
    bias = torch.cat([head[6].bias.detach().float(), head[6].bias.new_zeros(5)]).contiguous()   # epilogue reads 8-channel groups
    t = _conv_cl(t, head[6].weight, 1, None, bias)                      # (B,S,H,W,3)
    new_alpha = torch.empty((B, 3, S), dtype=torch.float32, device=dev)
    _rt.check(l.dff_spatial_mean_accum(_p(t), 3, B, S, H, W, _p(alpha), 0.001, 1.0, 1.0, _p(new_alpha), dev.index, _st(dev)))
    return new_alpha


def flow_forward(net, FS, FOVs):
    """`FlowNetwork.forward` (reference :63-104): the aligned focal stack (B,3,S,H,W)."""
    _rt._require_cuda(FS, "FS")
    _rt._require_cuda(FOVs, "FOVs")
    if net.training:
        raise _rt.DffError("dff_b200: the alignment network runs in eval mode only in this build")
    if FS.dim() != 5 or FS.shape[1] != 3 or FS.dtype != torch.float32:
        raise _rt.DffError("dff_b200: FS must be float32 (B,3,S,H,W), got %s %s" % (tuple(FS.shape), FS.dtype))
    B, _, S, H, W = FS.shape
    if S != 10:
        raise _rt.DffError("dff_b200: the alignment network is built for 10-slice stacks (AdaptiveAvgPool3d((10,1,1)), reference :40)")
    if H % 4 or W % 4:
        raise _rt.DffError("dff_b200: H and W must be multiples of 4 for the alignment network")
    _rt._check_device(FS.device.index)
    fov = FOVs.reshape(B, S).float().contiguous()
    with torch.no_grad():
        x = _rt.to_channels_last(FS, 4, False)
        fe1 = _block(net.OF_feature[1], _block(net.OF_feature[0], x, cin_pad_to=4))
        fe2 = _block(net.OF_feature1[1], _block(net.OF_feature1[0], fe1))
        fe3 = _block(net.OF_feature2[1], _block(net.OF_feature2[0], fe2))
        alpha = _align_head(net.conv1, fe3, None, fov)
        alpha = _align_head(net.conv2, fe2, alpha, fov)
        alpha = _align_head(net.conv3, fe1, alpha, fov)
        out, _ = _rt.fov_warp(FS, alpha.reshape(B, 3, S, 1, 1), FOVs, want_flow=False)
    return out
