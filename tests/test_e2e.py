"""End-to-End variant (alignment network + depth network): drop-in surface on CPU, parity with the reference's golden outputs
(tests/golden/g4_e2e_synth.npz, produced by the unmodified End_to_End/End_to_End.py) and the fp64 oracle on the GPU."""
import numpy as np
import pytest
import torch

from conftest import golden


def _net():
    from dffinthewild_b200.End_to_End import Network
    torch.manual_seed(0)
    return Network()


def test_e2e_state_dict_layout_and_seeded_init():
    lay = golden("state_layout_e2e.npz")
    sd = _net().state_dict()
    assert list(sd.keys()) == [str(k) for k in lay["keys"]]
    assert len(sd) == 522
    for (k, v), shp, s in zip(sd.items(), lay["shapes"], lay["seed0_sum"]):
        assert str(tuple(v.shape)) == str(shp), k
        assert abs(float(v.double().sum()) - s) <= 1e-9 * max(1.0, abs(s)) + 1e-9 * float(v.double().abs().sum()), k


def test_e2e_refuses_cpu_and_wrong_slice_count(built_lib):
    from dffinthewild_b200.runtime import DffError
    net = _net().eval()
    with pytest.raises(DffError):
        net(torch.zeros(1, 3, 10, 32, 32), torch.zeros(1, 10, 1, 1), torch.ones(1, 1, 10, 1, 1))


@pytest.mark.gpu
def test_e2e_matches_reference_golden(built_lib):
    from dffinthewild_b200 import synth
    g = golden("g4_e2e_synth.npz")
    net = _net()
    sd = synth.synthetic_state(net.state_dict(), seed=2)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    FS, fd = synth.focal_stack(1, 10, 32, 64, seed=14), synth.focus_dists(1, 10, 32, 64, "ddff", tiled=False)
    fov = synth.fovs(1, 10)
    with torch.no_grad():
        outs = net(FS.cuda(), fd.cuda(), fov.cuda())
    assert len(outs) == 5 and outs[4].shape == (1, 3, 10, 32, 64)
    assert np.abs(outs[4].cpu().numpy() - g["warped"]).max() <= 2e-4      # aligned stack (bilinear weights amplify ulp noise)
    for o, n in zip(outs[:4], ("mid", "p1", "p2", "p3")):
        ref = g[n]
        assert float((np.abs(o.cpu().numpy() - ref) / np.abs(ref)).max()) <= 1e-4, n


@pytest.mark.gpu
def test_e2e_alignment_vs_fp64_oracle_batch2(built_lib):
    """B = 2 exercises the reference's broadcast quirk (sample 0's scale correction applied to every sample)."""
    from oracle import dff_oracle as O
    from dffinthewild_b200 import synth
    net = _net()
    sd = synth.synthetic_state(net.state_dict(), seed=2)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    B, S, H, W = 2, 10, 32, 48
    FS = synth.focal_stack(B, S, H, W, seed=15)
    fov = synth.fovs(B, S) + 0.004 * torch.arange(B).view(B, 1, 1, 1, 1)
    with torch.no_grad():
        got = net.optical_flow_aggregation(FS.cuda(), fov.cuda()).cpu()
        ref = O.flow_forward({k: v.double() if v.is_floating_point() else v for k, v in sd.items()}, FS.double(), fov.double())
    assert (got.double() - ref).abs().max().item() <= 3e-4


def _e2e_net(precision="fp32"):
    from dffinthewild_b200 import synth
    net = _net()
    sd = synth.synthetic_state(net.state_dict(), seed=2)
    net.load_state_dict(sd, strict=True)
    net.optical_flow_aggregation.precision = precision
    net.DFF_net.precision = precision
    return net.cuda().eval(), sd


@pytest.mark.gpu
def test_alignment_bf16_tensor_core_path(built_lib):
    """bf16 mode of the alignment network (tcgen05 kernels on bf16 feature volumes; alpha, the final warp and the stack stay fp32)
    against the fp64 oracle.  What the network outputs is a sub-pixel similarity warp per slice, so the gate is on the warp
    parameters — shifts within 0.05 px, scale correction within 1e-4 (0.05 px at the border of a 1000 px image) — and on the
    aligned stack (mean |error| of a U(-1,1) noise image warped by a 0.05 px different shift is ~0.03)."""
    from oracle import dff_oracle as O
    from dffinthewild_b200 import synth
    from dffinthewild_b200.End_to_End import flow_forward
    net, sd = _e2e_net("bf16")
    B, S, H, W = 2, 10, 64, 96
    FS = synth.focal_stack(B, S, H, W, seed=17)
    fov = synth.fovs(B, S) + 0.004 * torch.arange(B).view(B, 1, 1, 1, 1)
    with torch.no_grad():
        got, alpha = flow_forward(net.optical_flow_aggregation, FS.cuda(), fov.cuda(), return_alpha=True)
        ref, ralpha = O.flow_forward({k: v.double() if v.is_floating_point() else v for k, v in sd.items()}, FS.double(), fov.double(),
                                     return_alpha=True)
    da = (alpha.cpu().double() - ralpha.reshape(B, 3, S)).abs()
    print("bf16 alignment: max |d scale| %.3g, max |d shift| %.3g px, mean |d stack| %.3g" % (
        da[:, 0].max().item(), da[:, 1:].max().item(), (got.cpu().double() - ref).abs().mean().item()))
    assert da[:, 0].max().item() <= 1e-4
    assert da[:, 1:].max().item() <= 0.05
    assert (got.cpu().double() - ref).abs().mean().item() <= 0.03


@pytest.mark.gpu
def test_e2e_c4_shape(built_lib):
    """BASELINE.json configs[3]: alignment + depth on one 10-slice 3x512x768 stack.  fp32: aligned stack against the
    fp64 oracle; bf16: warp parameters against the fp32 path and the four depth maps on the reference's metrics; CUDA-event time
    of both modes' alignment for profiles/."""
    from oracle import dff_oracle as O
    from dffinthewild_b200 import synth
    from dffinthewild_b200.End_to_End import flow_forward
    B, S, H, W = 1, 10, 512, 768
    FS = synth.focal_stack(B, S, H, W, seed=18)
    fd = synth.focus_dists(B, S, H, W, "ddff", tiled=False)
    fov = synth.fovs(B, S)
    net, sd = _e2e_net("fp32")
    with torch.no_grad():
        a32, alpha32 = flow_forward(net.optical_flow_aggregation, FS.cuda(), fov.cuda(), return_alpha=True)
        ref, ralpha = O.flow_forward({k: v.double() if v.is_floating_point() else v for k, v in sd.items()}, FS.double(), fov.double(),
                                     return_alpha=True)
    assert (a32.cpu().double() - ref).abs().max().item() <= 5e-4
    assert (alpha32.cpu().double() - ralpha.reshape(B, 3, S)).abs().max().item() <= 1e-4
    times = {}
    for prec in ("fp32", "bf16"):
        net, _ = _e2e_net(prec)
        with torch.no_grad():
            for _ in range(2):
                flow_forward(net.optical_flow_aggregation, FS.cuda(), fov.cuda())
            x = FS.cuda()
            f = fov.cuda()
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            for _ in range(5):
                al = flow_forward(net.optical_flow_aggregation, x, f)
            e1.record()
            for _ in range(5):
                outs = net.DFF_net(al, fd.cuda())
            e2.record()
            torch.cuda.synchronize()
            times[prec] = (e0.elapsed_time(e1) / 5, e1.elapsed_time(e2) / 5)
            if prec == "bf16":
                _, alpha16 = flow_forward(net.optical_flow_aggregation, x, f, return_alpha=True)
                full = net(x, fd.cuda(), f)
    d = (alpha16 - alpha32).abs()
    assert d[:, 0].max().item() <= 1e-4 and d[:, 1:].max().item() <= 0.05, d.amax(dim=(0, 2))
    assert len(full) == 5 and all(torch.isfinite(t).all() for t in full)
    import json
    import os
    dlog = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(dlog):
        with open(os.path.join(dlog, "parity_fullsize.jsonl"), "a") as fh:
            fh.write(json.dumps({"test": "e2e_c4", "alignment_ms": {k: v[0] for k, v in times.items()},
                                 "depth_ms": {k: v[1] for k, v in times.items()},
                                 "bf16_vs_fp32_alpha_max": [float(x) for x in d.amax(dim=(0, 2))]}) + "\n")
