"""Whole-network parity on the GPU through the drop-in module (which calls the C-ABI `dff_forward`):
against the reference's committed golden outputs and against the fp64 CPU oracle on seeded inputs."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu
NAMES = ("mid", "p1", "p2", "p3")
FP32_RTOL = 1e-4   # north_star: fp32 mode within 1e-4 relative per pixel


def _net(sd=None, precision="fp32"):
    from dffinthewild_b200.Depth_Estimation_Network import Network
    torch.manual_seed(0)
    net = Network()
    if sd is not None:
        net.load_state_dict(sd, strict=True)
    net.DFF_net.precision = precision
    return net.cuda().eval()


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float((np.abs(a - b) / np.abs(b)).max())


def test_golden_g1_asbuilt_fp32(built_lib):
    from dffinthewild_b200 import synth
    g = golden("g1_eval_asbuilt.npz")
    net = _net()
    FS, fd = synth.focal_stack(1, 3, 32, 64, seed=11), synth.focus_dists(1, 3, 32, 64, "ddff")
    from dffinthewild_b200 import runtime as rt
    with torch.no_grad():
        outs, costs = rt.dff_net_forward(net.DFF_net, FS.cuda(), fd.cuda(), return_costs=True)
    for c, n in zip(costs, ("cost_mid", "cost1", "cost2", "cost3")):
        ref = g[n]
        assert np.abs(c.cpu().numpy() - ref).max() <= 2e-5 * np.abs(ref).max(), n
    for o, n in zip(outs, NAMES):
        assert _rel(o.cpu().numpy(), g[n]) <= FP32_RTOL, n


def test_golden_g2_synth_fp32(built_lib):
    from dffinthewild_b200 import synth
    g = golden("g2_eval_synth.npz")
    net0 = _net()
    sd = synth.synthetic_state(net0.state_dict(), seed=1)
    net = _net(sd)
    FS, fd = synth.focal_stack(2, 5, 64, 32, seed=12, valid_hw=(60, 29)), synth.focus_dists(2, 5, 64, 32, "defocus")
    with torch.no_grad():
        outs = net(FS.cuda(), fd.cuda())
    for o, n in zip(outs, NAMES):
        assert o.shape == (2, 64, 32) and o.dtype == torch.float32 and o.is_cuda
        assert _rel(o.cpu().numpy(), g[n]) <= FP32_RTOL, n


@pytest.mark.parametrize("B,S,H,W,kind", [(1, 10, 96, 128, "ddff"), (3, 1, 32, 32, "defocus"), (1, 15, 64, 96, "ddff")])
def test_fp32_vs_fp64_oracle(built_lib, B, S, H, W, kind):
    from oracle import dff_oracle as O
    from dffinthewild_b200 import synth
    sd = synth.synthetic_state(_net().state_dict(), seed=4)
    net = _net(sd)
    FS = synth.focal_stack(B, S, H, W, seed=40 + S)
    fd = synth.focus_dists(B, S, H, W, kind, tiled=(S % 2 == 0))
    with torch.no_grad():
        outs = net(FS.cuda(), fd.cuda())
        ref = O.dff_forward({k: v.double() if v.is_floating_point() else v for k, v in sd.items()}, FS.double(), fd.double())
    for o, r, n in zip(outs, ref, NAMES):
        assert _rel(o.cpu().numpy(), r.numpy()) <= FP32_RTOL, n


def test_bf16_mode_on_reference_metrics(built_lib):
    """bf16 mode: tolerance on the reference's own depth metrics (metrics.py:90-97, 41-61), oracle as ground truth.
    Gates from SURVEY.md §7.3 for trained-like (calibrated) weights: AbsRel <= 1e-2, MSE <= 3e-6 x (range/0.26)^2."""
    from oracle import dff_oracle as O
    from dffinthewild_b200 import synth
    sd = synth.synthetic_state(_net().state_dict(), seed=4)
    net = _net(sd, "bf16")
    B, S, H, W = 1, 10, 96, 128
    FS, fd = synth.focal_stack(B, S, H, W, seed=50), synth.focus_dists(B, S, H, W, "ddff")
    with torch.no_grad():
        outs = net(FS.cuda(), fd.cuda())
        ref = O.dff_forward(sd, FS, fd)
    mask = np.ones((H, W), dtype=bool)
    for o, r, n in zip(outs, ref, NAMES):
        est, gt = o[0].cpu().numpy(), r[0].numpy()
        assert O.mask_abs_rel(est, gt, mask) <= 1e-2, n
        assert O.mask_mse(est, gt, mask) <= 3e-6, n
        assert O.bumpiness(gt, est, mask) <= 0.5, n


def test_weight_cache_follows_parameter_updates(built_lib):
    from dffinthewild_b200 import synth
    net = _net(synth.synthetic_state(_net().state_dict(), seed=4))
    FS, fd = synth.focal_stack(1, 2, 32, 32, seed=60).cuda(), synth.focus_dists(1, 2, 32, 32, "defocus").cuda()
    with torch.no_grad():
        a = net(FS, fd)[3].clone()
        b = net(FS, fd)[3].clone()
        assert torch.equal(a, b)
        net.DFF_net.classif3[0].weight.mul_(-1.0)
        c = net(FS, fd)[3]
    assert not torch.equal(a, c)


def test_rejects_bad_shapes(built_lib):
    from dffinthewild_b200.runtime import DffError
    net = _net()
    with pytest.raises(DffError):
        net(torch.zeros(1, 3, 2, 48, 64).cuda(), torch.zeros(1, 2, 48, 64).cuda())
    with pytest.raises(DffError):
        net(torch.zeros(1, 3, 2, 32, 32).cuda().double(), torch.zeros(1, 2, 32, 32).cuda())
