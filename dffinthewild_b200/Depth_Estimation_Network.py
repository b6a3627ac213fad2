"""Drop-in `Depth_Estimation_Network.Network` whose forward/backward run on hand-written sm_100a CUDA.

This file mirrors the *interface* of the reference module
(`train_codes/Depth_Estimation_Network.py:8-15` and the inference copy
`Depth_Estimation_Test/Depth_Estimation_Network.py:7-13`):

* `Network()` takes no arguments, has one child `DFF_net`;
* `forward(FS, focus_dists)` returns `(mid_out, pred1, pred2, pred3)`, each `(B, H, W)` fp32;
* the `state_dict` has exactly the reference's 384 keys / shapes / dtypes, so checkpoints move both ways;
* `torch.manual_seed(k); Network()` consumes the RNG stream in the reference's order (default construction of
  every conv / transposed conv in attribute order, then the N(0, sqrt(2/(k*Cout))) re-draw over `modules()`,
  reference lines 61-75), so as-built weights are identical.

The submodules below are *parameter containers only*: they are never called.  `forward` hands raw device
pointers to the C-ABI library (`include/dff_b200.h`).  There is no CPU path and no eager/PyTorch fallback:
a CPU tensor, a missing library or a non-sm_100 device raises.
"""
import math

import torch
import torch.nn as nn

from . import runtime as _rt

# ----------------------------------------------------------------------------------------------------------
# parameter containers (same attribute names / nesting / Sequential indices as the reference)
# ----------------------------------------------------------------------------------------------------------


def _conv_bn(cin, cout, k, stride, pad, dil=1):
    """`convbn_3d` (reference :352-355): bias-free Conv3d followed by BatchNorm3d, as Sequential[0], [1]."""
    return nn.Sequential(
        nn.Conv3d(cin, cout, kernel_size=k, stride=stride, padding=pad, dilation=dil, bias=False),
        nn.BatchNorm3d(cout),
    )


def _up_bn(cin, cout):
    """Exact-2x (H, W) transposed conv + BN (reference :43-50, 229-235, 296-301)."""
    return nn.Sequential(
        nn.ConvTranspose3d(cin, cout, kernel_size=3, padding=1, output_padding=(0, 1, 1), stride=(1, 2, 2), bias=False),
        nn.BatchNorm3d(cout),
    )


def _relu():
    return nn.ReLU(inplace=True)


class _Container(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - containers are never executed
        raise RuntimeError("parameter container: the network runs through the dff_b200 CUDA library, not nn.Module calls")


class _ResBlock2d(_Container):
    """`resnet_block_2d` (reference :361-370): two per-slice 1x3x3 conv+BN, ReLU between, residual, ReLU."""

    def __init__(self, c):
        super().__init__()
        self.conv = nn.Sequential(
            _conv_bn(c, c, (1, 3, 3), (1, 1, 1), (0, 1, 1), (1, 1, 1)),
            _relu(),
            _conv_bn(c, c, (1, 3, 3), (1, 1, 1), (0, 1, 1), (1, 1, 1)),
        )
        self.Relu = _relu()


class _SRD(_Container):
    """`Feature_Extraction` / `SRD` (reference :394-407 / test copy :317-330)."""

    def __init__(self, c):
        super().__init__()
        self.Focus_Measure = _ResBlock2d(c)
        self.N_ch_attention = nn.Sequential(
            nn.Conv3d(c, c, (3, 1, 1), stride=1, padding=(1, 0, 0), bias=False),
            _relu(),
            nn.Conv3d(c, c, 1, stride=1, padding=0, bias=False),
            _relu(),
        )


class _EFD(_Container):
    """`res_stride_conv_3d` / `EFD` (reference :383-392 / test copy :306-315)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.stride_conv = _conv_bn(cin, cout, (3, 3, 3), (1, 2, 2), (1, 1, 1))
        self.max_pooling = nn.Sequential(nn.MaxPool3d((1, 2, 2), (1, 2, 2)), _conv_bn(cin, cout, (3, 3, 3), 1, (1, 1, 1)))
        self.RELU = _relu()


class _FM(_Container):
    """`FM_module` (reference :141-153): 1x9x9 dilated focus-measure conv + SRD(8)."""

    def __init__(self):
        super().__init__()
        self.Focus_extraction = nn.Sequential(
            _conv_bn(3, 8, (1, 9, 9), 1, (0, 8, 8), (1, 2, 2)),
            _relu(),
            _SRD(8),
        )


class _Pyramid(_Container):
    """`hourglassup` (reference :179-273)."""

    def __init__(self, c):
        super().__init__()
        self.pooling_32 = nn.AvgPool3d((1, 8, 8), stride=(1, 8, 8))
        self.pooling_16 = nn.AvgPool3d((1, 4, 4), stride=(1, 4, 4))
        self.pooling_8 = nn.AvgPool3d((1, 2, 2), stride=(1, 2, 2))

        def tower(ci, co, last_relu):
            mods = [_conv_bn(ci, co, 3, 1, 1), _relu(), _conv_bn(co, co, 3, 1, 1)]
            if last_relu:
                mods.append(_relu())
            return nn.Sequential(*mods)

        self.dres8_0 = tower(c, c, True)
        self.dres8_1 = tower(c, c, False)
        self.dres16_0 = tower(c, 2 * c, True)
        self.dres16_1 = tower(2 * c, 2 * c, False)
        self.dres32_0 = tower(c, 2 * c, True)
        self.dres32_1 = tower(2 * c, 2 * c, False)
        self.conv1 = nn.Conv3d(c, 2 * c, kernel_size=3, stride=(1, 2, 2), padding=1, bias=False)
        self.conv2 = nn.Sequential(_conv_bn(2 * c, 2 * c, 3, 1, 1), _relu())
        self.conv3 = nn.Conv3d(2 * c, 4 * c, kernel_size=3, stride=(1, 2, 2), padding=1, bias=False)
        self.conv4 = nn.Sequential(_conv_bn(4 * c, 4 * c, 3, 1, 1), _relu())
        self.conv8 = _up_bn(4 * c, 2 * c)
        self.conv9 = _up_bn(2 * c, c)
        self.combine1 = nn.Sequential(_conv_bn(4 * c, 2 * c, 3, 1, 1), _relu())
        self.combine2 = nn.Sequential(_conv_bn(6 * c, 4 * c, 3, 1, 1), _relu())
        self.redir1 = _conv_bn(c, c, 1, 1, 0)
        self.redir2 = _conv_bn(2 * c, 2 * c, 1, 1, 0)
        self.redir3 = _conv_bn(4 * c, 4 * c, 1, 1, 0)  # defined, never executed (reference :244)


class _Hourglass(_Container):
    """`hourglass` (reference :275-321)."""

    def __init__(self, p):
        super().__init__()
        self.conv0 = nn.Sequential(_conv_bn(2 * p, p, 3, 1, (1, 1, 1)), _relu())
        self.conv1 = nn.Sequential(_conv_bn(p, 2 * p, 3, (1, 2, 2), (1, 1, 1)), _relu())
        self.pre_conv = nn.Sequential(_conv_bn(2 * p, 2 * p, 1, 1, 0), _relu())  # never executed (reference :285-286)
        self.conv2 = _conv_bn(2 * p, 2 * p, 3, 1, 1)
        self.conv3 = nn.Sequential(_conv_bn(2 * p, 2 * p, 3, (1, 2, 2), 1), _relu())
        self.conv4 = nn.Sequential(_conv_bn(2 * p, 2 * p, 3, 1, 1), _relu())
        self.conv5 = _up_bn(2 * p, 2 * p)
        self.conv6 = _up_bn(2 * p, p)


def reference_init_(net):
    """The reference's post-construction initialisation (reference :61-75), applied over `modules()` order.

    `nn.ConvTranspose3d` is not an `nn.Conv3d`, so transposed convs keep PyTorch's default init.
    """
    for m in net.modules():
        if isinstance(m, nn.Conv3d):
            n = m.kernel_size[0] * m.kernel_size[1] * m.kernel_size[2] * m.out_channels
            m.weight.data.normal_(0, math.sqrt(2.0 / n))
        elif isinstance(m, nn.BatchNorm3d):
            m.weight.data.fill_(1)
            m.bias.data.zero_()


class DFF_net(_rt.PackedOwnerMixin, nn.Module):
    """Depth-from-focus network (reference `DFF_net`, :17-137).  Forward = one C-ABI call."""

    def __init__(self):
        super().__init__()
        self.FM_measure = _FM()
        self.FM_conv1 = nn.Sequential(_EFD(8, 16), _SRD(16))
        self.FM_conv2 = nn.Sequential(_EFD(16, 32), _SRD(32))
        self.SPP_module = _Pyramid(32)
        self.confidence = nn.Sequential(_conv_bn(32, 32, 3, 1, 1), _relu(),
                                        nn.Conv3d(32, 1, kernel_size=3, padding=1, stride=1, bias=False))
        self.dres0 = nn.Sequential(_conv_bn(32, 64, 3, 1, 1), _relu(), _conv_bn(64, 64, 3, 1, 1), _relu())
        self.deconv_1 = _up_bn(64, 32)
        self.dres2 = _Hourglass(32)
        self.deconv_2 = _up_bn(32, 16)
        self.dres3 = _Hourglass(16)
        self.deconv_3 = _up_bn(16, 8)
        self.dres4 = _Hourglass(8)
        self.classif1 = nn.Sequential(nn.Conv3d(32, 1, kernel_size=1, padding=0, stride=1, bias=False))
        self.classif2 = nn.Sequential(nn.Conv3d(16, 1, kernel_size=1, padding=0, stride=1, bias=False))
        self.classif3 = nn.Sequential(nn.Conv3d(8, 1, kernel_size=1, padding=0, stride=1, bias=False))
        reference_init_(self)
        # precision of the CUDA path: "fp32" (parity mode, FFMA, <=1e-4 rel/pixel) or "bf16" (tcgen05 throughput mode)
        self.precision = _rt.default_precision()

    def forward(self, FS, focus_dists):
        return _rt.dff_net_forward(self, FS, focus_dists)


class Network(nn.Module):
    """Reference `Network` (:8-15): delegation to `DFF_net`."""

    def __init__(self):
        super().__init__()
        self.DFF_net = DFF_net()

    def forward(self, FS, focus_dists):
        return self.DFF_net(FS, focus_dists)
