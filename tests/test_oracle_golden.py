"""The CPU oracle against the committed reference outputs (tests/golden, made by oracle/gen_golden.py from the
unmodified reference).  Same torch build -> bit-exact; a different build may reorder float sums, so the gate is a
tight relative tolerance rather than equality."""
import numpy as np
import torch

from conftest import golden
from oracle import dff_oracle as O
from dffinthewild_b200 import synth

RTOL = 2e-5


def _template():
    from dffinthewild_b200.Depth_Estimation_Network import Network
    torch.manual_seed(0)
    return Network().state_dict()


def _close(a, b, tol=RTOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    err = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
    assert err.max() <= tol, err.max()


def test_g1_eval_asbuilt():
    g = golden("g1_eval_asbuilt.npz")
    sd = _template()
    FS, fd = synth.focal_stack(1, 3, 32, 64, seed=11), synth.focus_dists(1, 3, 32, 64, "ddff")
    rec = {}
    with torch.no_grad():
        o = O.dff_forward(sd, FS, fd, record=rec)
    for t, n in zip(o, ("mid", "p1", "p2", "p3")):
        _close(t, g[n])
    for n in ("cost_mid", "cost1", "cost2", "cost3"):
        _close(rec[n], g[n], 1e-4)


def test_g2_eval_synth():
    g = golden("g2_eval_synth.npz")
    sd = synth.synthetic_state(_template(), seed=1)
    FS, fd = synth.focal_stack(2, 5, 64, 32, seed=12, valid_hw=(60, 29)), synth.focus_dists(2, 5, 64, 32, "defocus")
    rec = {}
    with torch.no_grad():
        o = O.dff_forward(sd, FS, fd, record=rec)
    for t, n in zip(o, ("mid", "p1", "p2", "p3")):
        _close(t, g[n])
    assert abs(float(rec["V3"].double().sum()) - float(g["V3_sum"])) <= 1e-6 * abs(float(g["V3_sum"])) + 1e-3


def test_g3_train_loss_and_grads():
    g = golden("g3_train_synth.npz")
    sd = synth.synthetic_state(_template(), seed=1)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    FS, fd = synth.focal_stack(2, 4, 32, 32, seed=13), synth.focus_dists(2, 4, 32, 32, "defocus")
    gt, mask = synth.gt_and_mask(2, 32, 32, seed=13)
    o = O.dff_forward(sd, FS, fd, train=True)
    loss = O.defocus_loss(o, gt, mask)
    loss.backward()
    _close(loss.detach(), g["loss"], 1e-5)
    names = [str(n) for n in g["grad_names"]]
    for n, s_, a_ in zip(names, g["grad_sum"], g["grad_abs"]):
        gr = sd[n].grad
        assert gr is not None, n
        assert abs(float(gr.double().abs().sum()) - a_) <= 1e-3 * a_ + 1e-9, n
    for n in g["grad_none"]:
        assert sd[str(n)].grad is None


def test_g4_e2e():
    g = golden("g4_e2e_synth.npz")
    lay = golden("state_layout_e2e.npz")
    tmpl = {str(k): torch.zeros(eval(str(s)), dtype=torch.int64 if str(k).endswith("num_batches_tracked") else torch.float32)
            for k, s in zip(lay["keys"], lay["shapes"])}
    sd = synth.synthetic_state(tmpl, seed=2)
    FS, fd = synth.focal_stack(1, 10, 32, 64, seed=14), synth.focus_dists(1, 10, 32, 64, "ddff", tiled=False)
    with torch.no_grad():
        o = O.e2e_forward(sd, FS, fd, synth.fovs(1, 10))
    for t, n in zip(o, ("mid", "p1", "p2", "p3", "warped")):
        _close(t, g[n], 1e-4)


def test_g5_fov_warp():
    for tag in ("b1", "b2"):
        g = golden("g5_fov_warp_%s.npz" % tag)
        out, flow = O.fov_warp(torch.from_numpy(g["x"]), torch.from_numpy(g["alpha"]), torch.from_numpy(g["fov"]))
        _close(flow, g["flow"], 1e-5)
        assert np.abs(out.numpy() - g["out"]).max() < 1e-5


def test_metrics_restatement():
    rng = np.random.default_rng(0)
    gt = rng.uniform(0.1, 1.0, (24, 32))
    est = gt + rng.normal(0, 0.01, gt.shape)
    m = np.ones_like(gt, dtype=bool)
    assert abs(O.mask_mse(est, gt, m) - np.mean((est - gt) ** 2)) < 1e-15
    assert O.mask_abs_rel(est, gt, m) > 0
    assert 0 < O.bumpiness(gt, est, m) <= 5.0
    assert O.bumpiness(gt, gt, m) == 0.0


def test_g7_three_training_steps():
    """Three steps of the reference's training loop (train_code_Defocus.py:67,158-168): oracle forward + torch.optim.Adam on the
    state_dict tensors reproduces the reference module's loss trajectory and final parameters."""
    g = golden("g7_train_3steps.npz")
    sd = synth.synthetic_state(_template(), seed=1)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    opt = torch.optim.Adam([v for v in sd.values() if v.requires_grad], lr=1e-3, betas=(0.9, 0.99))
    for step in range(3):
        FS, fd = synth.focal_stack(2, 4, 32, 32, seed=20 + step), synth.focus_dists(2, 4, 32, 32, "defocus")
        gt, mask = synth.gt_and_mask(2, 32, 32, seed=20 + step)
        o = O.dff_forward(sd, FS, fd, train=True, update_stats=True)
        opt.zero_grad()
        loss = O.defocus_loss(o, gt, mask)
        loss.backward()
        opt.step()
        _close(loss.detach(), g["losses"][step], 1e-4)
    for k in ("DFF_net.classif3.0.weight", "DFF_net.dres4.conv6.1.weight", "DFF_net.dres4.conv6.1.running_var"):
        _close(sd[k].detach(), g["w:" + k], 2e-3)


def test_g8_metrics_restatement_matches_reference_metrics_py():
    """oracle.depth_metrics against the values the reference's own metrics.py (:90-133) produced (tests/golden/g8_metrics.npz)."""
    g = golden("g8_metrics.npz")
    for b in range(2):
        o = O.depth_metrics(g["est"][b], g["gt"][b], g["mask"][b], g["conf"][b])
        for k, v in o.items():
            assert abs(v - g[k][b]) <= 1e-6 * abs(g[k][b]), (k, v, g[k][b])
