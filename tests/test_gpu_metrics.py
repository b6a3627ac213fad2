"""On-device evaluation tail: the reference's masked metrics (metrics.py:90-133) and the crop + colour map of
Depth_Estimation_Test/test.py:123-140, against the reference's own values (tests/golden/g8_metrics.npz) and the oracle."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu
KEYS = ("abs_rel", "sq_rel", "mse", "mae", "rmse", "rmse_log", "accuracy_1", "accuracy_2", "accuracy_3", "mse_w_conf", "mae_w_conf")


def test_metrics_match_reference_values(built_lib):
    from dffinthewild_b200 import metrics as M
    g = golden("g8_metrics.npz")
    est, gt = torch.from_numpy(g["est"]), torch.from_numpy(g["gt"])
    mask, conf = torch.from_numpy(g["mask"]), torch.from_numpy(g["conf"])
    # the network returns maps padded to multiples of 32: metrics must ignore the padding (test.py:125 crops first)
    padded = torch.full((2, 64, 96), 7.0)
    padded[:, :61, :90] = est
    fig = M.depth_metrics(padded.cuda(), gt.cuda(), mask.cuda(), conf.cuda())
    for k in KEYS:
        got = fig[k].cpu().numpy().astype(np.float64)
        tol = 0.0 if k.startswith("accuracy") else 2e-6     # the reference sums in fp32, the kernel in fp64
        assert np.all(np.abs(got - g[k]) <= tol * np.abs(g[k]) + (1e-7 if tol == 0.0 else 0.0)), (k, got, g[k])
    assert fig["count"].cpu().tolist() == [float(mask[b].sum()) for b in range(2)]
    # the reference's one-map signatures
    assert abs(M.mask_abs_rel(est[0].cuda(), gt[0].cuda(), mask[0].cuda()).item() - g["abs_rel"][0]) <= 2e-6 * g["abs_rel"][0]
    assert abs(M.mask_mse_w_conf(est[1].cuda(), gt[1].cuda(), conf[1].cuda(), mask[1].cuda()).item() - g["mse_w_conf"][1]) <= 2e-6 * g["mse_w_conf"][1]
    assert abs(M.mask_accuracy_k(est[0].cuda(), gt[0].cuda(), 1, mask[0].cuda()).item() - g["accuracy_1"][0]) <= 1e-7
    # no mask = all valid
    from oracle import dff_oracle as O
    ref = O.depth_metrics(g["est"][0], g["gt"][0], np.ones_like(g["mask"][0]))
    fig = M.depth_metrics(est[:1].cuda(), gt[:1].cuda())
    for k in KEYS[:9]:
        assert abs(fig[k][0].item() - ref[k]) <= 2e-6 * abs(ref[k]) + 1e-7, k


def test_jet_output_image(built_lib):
    """Crop + normalise + 'jet' + uint8 (test.py:123-140).  matplotlib is not installed, so this pins the published segment data of
    the colormap (parity with matplotlib's table is unpinned): end points, mid point, monotone hue, crop."""
    from dffinthewild_b200 import metrics as M
    H, W, Hc, Wc = 32, 64, 29, 61
    est = torch.linspace(0.02, 0.28, H * W).view(1, H, W).cuda()
    rgb = M.depth_to_jet(est, (Hc, Wc), 0.02, 0.28)
    assert rgb.shape == (1, Hc, Wc, 3) and rgb.dtype == torch.uint8
    lut = M.depth_to_jet(torch.linspace(0, 1, 256).view(1, 1, 256).cuda() * (255.0 / 256.0) + 0.5 / 256, (1, 256), 0.0, 1.0)[0, 0].cpu().numpy()
    assert tuple(lut[0]) == (0, 0, 128) and tuple(lut[255]) == (128, 0, 0)          # jet: dark blue ... dark red
    assert lut[127][1] == 255 and lut[127][0] < 140 and lut[127][2] < 140            # green plateau in the middle
    assert np.all(np.diff(lut[:, 0].astype(int)[:227]) >= 0)                        # red rises until its plateau ends
    # out-of-range values clip to the end colours; the crop reads the right pixels
    clip = M.depth_to_jet(torch.tensor([[[-5.0, 0.5, 9.0]]]).cuda(), (1, 3), 0.0, 1.0)[0, 0].cpu().numpy()
    assert tuple(clip[0]) == (0, 0, 128) and tuple(clip[2]) == (128, 0, 0)
    v = (est[0, 5, 7].item() - 0.02) / 0.26
    assert tuple(rgb[0, 5, 7].cpu().numpy()) == tuple(lut[min(255, int(v * 256))])
