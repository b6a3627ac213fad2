"""Generate tests/golden/*.npz from the UNMODIFIED reference — run in the build container only.

    python oracle/gen_golden.py            # needs /root/reference (read-only); writes tests/golden/

For every case it (1) imports the reference module by path, (2) loads deterministic weights, (3) runs the
reference, (4) runs `oracle/dff_oracle.py` on the same tensors and asserts the restatement is bit-identical to
the reference (this is what pins the oracle), and (5) stores the reference outputs as small fp32 fixtures.
The reference cannot travel to the GPU box; the fixtures and this script do.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dff_oracle as O  # noqa: E402
from dffinthewild_b200 import synth  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def same(a, b, what):
    assert a.shape == b.shape and torch.equal(a, b), "oracle != reference at %s (max diff %g)" % (
        what, (a - b).abs().max().item())


def new_cases(tden, only_g8=False):
    """Round-2 additions (kept separate so the round-1 fixtures are not rewritten): G6 = train step at the C3 shape, G7 = three
    optimizer steps of the reference's training loop."""
    if not only_g8:
        _train_cases(tden)
    _metric_cases()


def _train_cases(tden):
    torch.manual_seed(0)
    ref = tden.Network()
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    sd1 = synth.synthetic_state(sd0, seed=1)
    crit = torch.nn.MSELoss()

    def defocus_loss(r, gt, mask):    # train_code_Defocus.py:160-165
        return 0.5 * crit(r[1][mask], gt[mask]) + 0.7 * crit(r[2][mask], gt[mask]) + 1.0 * crit(r[3][mask], gt[mask]) \
            + 0.3 * crit(r[0][mask], gt[mask])

    # ---- G6: train-mode forward + loss + gradients at BASELINE configs[2]'s stack shape: 2 x (5 slices, 256x256) -------------
    ref.load_state_dict(sd1, strict=True)
    ref.train()
    FS, fd = synth.focal_stack(2, 5, 256, 256, seed=16), synth.focus_dists(2, 5, 256, 256, "defocus")
    gt, mask = synth.gt_and_mask(2, 256, 256, seed=16)
    r = ref(FS, fd)
    loss = defocus_loss(r, gt, mask)
    loss.backward()
    sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd1.items()}
    o = O.dff_forward(sdo, FS, fd, train=True)
    lo = O.defocus_loss(o, gt, mask)
    lo.backward()
    for a, b, n in zip(r, o, ("mid", "p1", "p2", "p3")):
        same(a.detach(), b.detach(), "G6." + n)
    same(loss.detach(), lo.detach(), "G6.loss")
    gnames, gsum, gabs, gl2, small = [], [], [], [], {}
    for k, p in ref.named_parameters():
        if p.grad is None:
            assert sdo[k].grad is None, k
            continue
        same(p.grad, sdo[k].grad, "G6.grad." + k)
        gnames.append(k); gsum.append(float(p.grad.double().sum())); gabs.append(float(p.grad.double().abs().sum()))
        gl2.append(float(p.grad.double().norm()))
        if p.numel() <= 2400:
            small["grad:" + k] = p.grad.numpy().copy()
    new_sd = ref.state_dict()
    bn = {("bn:" + k): v.numpy().copy() for k, v in new_sd.items() if "running" in k and ("dres4" in k or "FM_measure" in k)}
    sub = lambda t: t.detach()[:, ::8, ::8].numpy().copy()
    np.savez_compressed(os.path.join(OUT, "g6_train_c3.npz"), mid=sub(r[0]), p1=sub(r[1]), p2=sub(r[2]), p3=sub(r[3]),
                        out_sum=np.array([float(t.detach().double().sum()) for t in r]), loss=loss.detach(),
                        grad_names=np.array(gnames), grad_sum=np.array(gsum), grad_abs=np.array(gabs), grad_l2=np.array(gl2),
                        **small, **bn)

    # ---- G7: three steps of the reference's training loop (train_code_Defocus.py:67,158-168), 2 x (4 slices, 32x32) --------
    ref.load_state_dict(sd1, strict=True)
    ref.train()
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3, betas=(0.9, 0.99))
    losses = []
    for step in range(3):
        FS, fd = synth.focal_stack(2, 4, 32, 32, seed=20 + step), synth.focus_dists(2, 4, 32, 32, "defocus")
        gt, mask = synth.gt_and_mask(2, 32, 32, seed=20 + step)
        r = ref(FS, fd)
        opt.zero_grad()
        loss = defocus_loss(r, gt, mask)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    fin = ref.state_dict()
    keys = [k for k in fin if fin[k].is_floating_point()]
    np.savez_compressed(os.path.join(OUT, "g7_train_3steps.npz"), losses=np.array(losses), keys=np.array(keys),
                        p_sum=np.array([float(fin[k].double().sum()) for k in keys]),
                        p_abs=np.array([float(fin[k].double().abs().sum()) for k in keys]),
                        **{("w:" + k): fin[k].numpy().copy() for k in ("DFF_net.classif3.0.weight", "DFF_net.dres4.conv6.1.weight",
                                                                       "DFF_net.dres4.conv6.1.running_var",
                                                                       "DFF_net.FM_measure.Focus_extraction.0.1.bias")})
    print("G6 loss", float(lo), "G7 losses", losses)


def _metric_cases():
    # ---- G8: the reference's metrics.py (imported unmodified behind a stub for the absent scikit-image, which only get_bumpiness uses)
    import types
    sk, skf = types.ModuleType("skimage"), types.ModuleType("skimage.filters")
    sk.filters = skf
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.filters", skf)
    met = load("Depth_Estimation_Test/metrics.py", "ref_metrics")
    rng = np.random.Generator(np.random.PCG64(88))
    gt = rng.uniform(0.05, 1.5, (2, 61, 90)).astype(np.float32)
    est = (gt * rng.uniform(0.7, 1.4, gt.shape)).astype(np.float32)
    mask = rng.uniform(0, 1, gt.shape) > 0.15
    conf = rng.uniform(0, 1, gt.shape).astype(np.float32)
    vals = {}
    for b in range(2):
        e, g, m, c = est[b], gt[b], mask[b], conf[b]
        r = {"abs_rel": met.mask_abs_rel(e, g, m), "sq_rel": met.mask_sq_rel(e, g, m), "mse": met.mask_mse(e, g, m),
             "mae": met.mask_mae(e, g, m), "rmse": met.mask_rmse(e, g, m), "rmse_log": met.mask_rmse_log(e, g, m),
             "accuracy_1": met.mask_accuracy_k(e, g, 1, m), "accuracy_2": met.mask_accuracy_k(e, g, 2, m),
             "accuracy_3": met.mask_accuracy_k(e, g, 3, m), "mse_w_conf": met.mask_mse_w_conf(e, g, c, m),
             "mae_w_conf": met.mask_mae_w_conf(e, g, c, m)}
        o = O.depth_metrics(e, g, m, c)
        for k, v in r.items():
            assert float(v) == o[k], ("G8", k, float(v), o[k])
            vals.setdefault(k, []).append(float(v))
    np.savez_compressed(os.path.join(OUT, "g8_metrics.npz"), est=est, gt=gt, mask=mask, conf=conf,
                        **{k: np.array(v, dtype=np.float64) for k, v in vals.items()})
    print("G8", {k: v[0] for k, v in vals.items()})


def main():
    torch.set_num_threads(8)
    torch.use_deterministic_algorithms(True)
    os.makedirs(OUT, exist_ok=True)
    tden = load("train_codes/Depth_Estimation_Network.py", "ref_tden")
    if "--new" in sys.argv:
        new_cases(tden, only_g8="--g8" in sys.argv)
        return
    eden = load("Depth_Estimation_Test/Depth_Estimation_Network.py", "ref_eden")
    e2e = load("End_to_End/End_to_End.py", "ref_e2e")

    # ---- state_dict layout + as-built checksums (seeded construction) -------------------------------------
    torch.manual_seed(0)
    ref = tden.Network()
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    keys = list(sd0.keys())
    np.savez_compressed(
        os.path.join(OUT, "state_layout.npz"),
        keys=np.array(keys), shapes=np.array([str(tuple(v.shape)) for v in sd0.values()]),
        dtypes=np.array([str(v.dtype) for v in sd0.values()]),
        seed0_sum=np.array([float(v.double().sum()) for v in sd0.values()]),
        seed0_abs=np.array([float(v.double().abs().sum()) for v in sd0.values()]))
    torch.manual_seed(0)
    sde = eden.Network().state_dict()
    assert list(sde.keys()) == keys and all(torch.equal(sde[k], sd0[k]) for k in keys)

    # ---- G1: eval, as-built weights (seed 0), 1 x (3 slices, 32x64) ----------------------------------------
    ref.eval()
    FS, fd = synth.focal_stack(1, 3, 32, 64, seed=11), synth.focus_dists(1, 3, 32, 64, "ddff")
    with torch.no_grad():
        r = ref(FS, fd)
        rec = {}
        o = O.dff_forward(sd0, FS, fd, record=rec)
    for a, b, n in zip(r, o, ("mid", "p1", "p2", "p3")):
        same(a, b, "G1." + n)
    np.savez_compressed(os.path.join(OUT, "g1_eval_asbuilt.npz"), mid=r[0], p1=r[1], p2=r[2], p3=r[3],
                        cost_mid=rec["cost_mid"], cost1=rec["cost1"], cost2=rec["cost2"], cost3=rec["cost3"])

    # ---- G2: eval, synthetic trained-like weights, 2 x (5 slices, 64x32), padded border ----------------------
    sd1 = synth.synthetic_state(sd0, seed=1)
    ref.load_state_dict(sd1, strict=True)
    FS, fd = synth.focal_stack(2, 5, 64, 32, seed=12, valid_hw=(60, 29)), synth.focus_dists(2, 5, 64, 32, "defocus")
    with torch.no_grad():
        r = ref(FS, fd)
        rec = {}
        o = O.dff_forward(sd1, FS, fd, record=rec)
        re_ = eden.Network(); re_.load_state_dict(sd1); re_.eval()
        r2 = re_(FS, fd)
    for a, b, b2, n in zip(r, o, r2, ("mid", "p1", "p2", "p3")):
        same(a, b, "G2." + n); same(a, b2, "G2.eden." + n)
    np.savez_compressed(os.path.join(OUT, "g2_eval_synth.npz"), mid=r[0], p1=r[1], p2=r[2], p3=r[3],
                        cost_mid=rec["cost_mid"], cost1=rec["cost1"], cost2=rec["cost2"], cost3=rec["cost3"],
                        V3_sum=np.array(float(rec["V3"].double().sum())),
                        FS_volume_sum=np.array(float(rec["FS_volume"].double().sum())))

    # ---- G3: train-mode forward + Defocus loss + gradients, 2 x (4 slices, 32x32) -------------------------
    ref.load_state_dict(sd1, strict=True)
    ref.train()
    FS, fd = synth.focal_stack(2, 4, 32, 32, seed=13), synth.focus_dists(2, 4, 32, 32, "defocus")
    gt, mask = synth.gt_and_mask(2, 32, 32, seed=13)
    r = ref(FS, fd)
    crit = torch.nn.MSELoss()
    loss = 0.5 * crit(r[1][mask], gt[mask]) + 0.7 * crit(r[2][mask], gt[mask]) + 1.0 * crit(r[3][mask], gt[mask]) \
        + 0.3 * crit(r[0][mask], gt[mask])                    # train_code_Defocus.py:160-165
    loss.backward()
    sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
           for k, v in sd1.items()}
    o = O.dff_forward(sdo, FS, fd, train=True)
    lo = O.defocus_loss(o, gt, mask)
    lo.backward()
    for a, b, n in zip(r, o, ("mid", "p1", "p2", "p3")):
        same(a.detach(), b.detach(), "G3." + n)
    same(loss.detach(), lo.detach(), "G3.loss")
    gnames, gsum, gabs, gnone = [], [], [], []
    small = {}
    for k, p in ref.named_parameters():
        if p.grad is None:
            gnone.append(k); assert sdo[k].grad is None, k
            continue
        same(p.grad, sdo[k].grad, "G3.grad." + k)
        gnames.append(k); gsum.append(float(p.grad.double().sum())); gabs.append(float(p.grad.double().abs().sum()))
        if p.numel() <= 4096:
            small["grad:" + k] = p.grad.numpy().copy()
    new_sd = ref.state_dict()
    bn = {("bn:" + k): v.numpy().copy() for k, v in new_sd.items() if "running" in k and "dres4" in k}
    np.savez_compressed(os.path.join(OUT, "g3_train_synth.npz"), mid=r[0].detach(), p1=r[1].detach(),
                        p2=r[2].detach(), p3=r[3].detach(), loss=loss.detach(), grad_names=np.array(gnames),
                        grad_sum=np.array(gsum), grad_abs=np.array(gabs), grad_none=np.array(gnone), **small, **bn)

    # ---- G4: End-to-End (alignment + depth), 1 x (10 slices, 32x64) ---------------------------------------
    torch.manual_seed(0)
    refe = e2e.Network()
    sde0 = {k: v.clone() for k, v in refe.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, "state_layout_e2e.npz"), keys=np.array(list(sde0.keys())),
                        shapes=np.array([str(tuple(v.shape)) for v in sde0.values()]),
                        seed0_sum=np.array([float(v.double().sum()) for v in sde0.values()]))
    sde1 = synth.synthetic_state(sde0, seed=2)
    refe.load_state_dict(sde1, strict=True)
    refe.eval()
    FS, fd = synth.focal_stack(1, 10, 32, 64, seed=14), synth.focus_dists(1, 10, 32, 64, "ddff", tiled=False)
    fov = synth.fovs(1, 10)
    with torch.no_grad():
        r = refe(FS, fd, fov)
        o = O.e2e_forward(sde1, FS, fd, fov)
    for i, (a, b) in enumerate(zip(r, o)):
        same(a, b, "G4.%d" % i)
    np.savez_compressed(os.path.join(OUT, "g4_e2e_synth.npz"), mid=r[0], p1=r[1], p2=r[2], p3=r[3], warped=r[4])

    # ---- G5: FOV_warp alone with non-trivial alpha, incl. the batch>1 broadcast quirk ---------------------
    flow_net = refe.optical_flow_aggregation
    g = torch.Generator().manual_seed(5)
    for B, tag in ((1, "b1"), (2, "b2")):
        x = torch.rand(B, 8, 10, 24, 40, generator=g) * 2 - 1
        alpha = torch.randn(B, 3, 10, 1, 1, generator=g) * torch.tensor([0.01, 1.5, 1.5]).view(1, 3, 1, 1, 1)
        fov = synth.fovs(B, 10) + 0.003 * torch.arange(B).view(B, 1, 1, 1, 1)
        with torch.no_grad():
            ro, rf = flow_net.FOV_warp(x, alpha, fov)
            oo, of = O.fov_warp(x, alpha, fov)
        same(ro, oo, "G5.out." + tag); same(rf, of, "G5.flow." + tag)
        np.savez_compressed(os.path.join(OUT, "g5_fov_warp_%s.npz" % tag), x=x, alpha=alpha, fov=fov, out=ro, flow=rf)
    print("golden fixtures written to", OUT)
    os.system("ls -la %s" % OUT)


if __name__ == "__main__":
    main()
