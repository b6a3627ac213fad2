"""Drop-in `End_to_End.Network` (reference End_to_End/End_to_End.py:8-17): the alignment network `FlowNetwork(8)` followed by
`DFF_net`, both on hand-written sm_100a CUDA through the C-ABI library.

Interface contract (reference `test_real_scenes.py:11,20-34`): `Network()` takes no arguments, has the children `DFF_net` and
`optical_flow_aggregation`; `forward(FS, focus_dists, FOVs)` returns `(mid_out, pred1, pred2, pred3, warped_FS)`; the
`state_dict` has the reference's 522 keys.  `torch.manual_seed(k); Network()` reproduces the reference's as-built weights
(`DFF_net` with its N(0, sqrt(2/n)) re-draw first, then `FlowNetwork` with PyTorch's default initialisation).

`FlowNetwork.forward` (reference :63-104) runs in eval mode only in this build (the reference ships no training script for it) as
ONE C-ABI call, `dff_flow_forward`: six `resnet_block_2d_OF` blocks and three alignment heads (BatchNorm folded into the conv
epilogues once, when the weights are packed), the FOV warps of the feature volumes, the pairwise (last slice ‖ slice ‖ flow) volumes,
the per-slice spatial means with the 0.001 factor and the running sum of alpha, and the final warp of the focal stack.  Precision
follows `FlowNetwork.precision`: "fp32" = FFMA parity path, "bf16" = tcgen05 kernels on bf16 feature volumes (alpha, the final warp
and the stack itself stay fp32).  Like the reference, it requires S == 10 and is bug-compatible for B > 1 (sample 0's scale
correction is applied to every sample).
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


class FlowNetwork(_rt.PackedOwnerMixin, nn.Module):
    """Alignment network (reference `FlowNetwork`, :18-104): parameter containers + a forward that is one C-ABI call."""

    def __init__(self, inplanes):
        super().__init__()
        self.precision = _rt.default_precision()   # "fp32" (FFMA parity path) or "bf16" (tcgen05 kernels, bf16 feature volumes)
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
# FlowNetwork.forward = ONE C-ABI call (dff_flow_forward); weights packed (BatchNorm folded) once per parameter version
# ---------------------------------------------------------------------------------------------------------------
def _p(t):
    return _P(t.data_ptr()) if t is not None else _P(0)


def _st(dev):
    return _P(torch.cuda.current_stream(dev).cuda_stream)


def _declare(l):
    if getattr(l, "_dff_flow_declared", False):
        return
    c = ctypes
    l.dff_flow_workspace_bytes.restype = c.c_size_t
    l.dff_flow_workspace_bytes.argtypes = [c.c_int] * 5
    l.dff_flow_forward.restype = c.c_int
    l.dff_flow_forward.argtypes = [_P, _P, _P] + [c.c_int] * 4 + [_P, _P, _P, c.c_size_t, c.c_int, c.c_int, _P]
    l._dff_flow_declared = True


_flow_ws = {}


def flow_forward(net, FS, FOVs, return_alpha=False):
    """`FlowNetwork.forward` (reference :63-104): the aligned focal stack (B,3,S,H,W) [and the estimated per-slice warp (B,3,S)]."""
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
    dev = FS.device
    _rt._check_device(dev.index)
    l = _rt.lib()
    _declare(l)
    mode = _rt._mode(net)
    fov = FOVs.reshape(B, S).float().contiguous()
    FS = FS.contiguous()
    packed = _rt.packed_weights(net, dev, _rt.NET_FLOW)
    n = l.dff_flow_workspace_bytes(B, S, H, W, mode)
    if n == 0:
        _rt.check(-1)
    with _rt._ws_lock:
        ws = _flow_ws.get(dev.index)
        if ws is None or ws.numel() < n:
            _flow_ws.pop(dev.index, None)
            ws = torch.empty(n, dtype=torch.uint8, device=dev)
            _flow_ws[dev.index] = ws
    out = torch.empty_like(FS)
    alpha = torch.empty((B, 3, S), dtype=torch.float32, device=dev) if return_alpha else None
    with torch.cuda.device(dev):
        _rt.check(l.dff_flow_forward(_p(packed), _p(FS), _p(fov), B, S, H, W, _p(out), _p(alpha), _p(ws), ws.numel(), mode, dev.index,
                                     _st(dev)))
    return (out, alpha) if return_alpha else out
