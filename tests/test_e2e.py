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
