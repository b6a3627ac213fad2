"""Train-mode parity on the GPU: batch-statistics BatchNorm forward, the Defocus loss recipe
(train_code_Defocus.py:160-165), backward through every operator, parameter gradients — against the reference's committed
golden gradients (tests/golden/g3_train_synth.npz, produced by the unmodified reference) and against the fp64 CPU oracle."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu
NAMES = ("mid", "p1", "p2", "p3")


def _net(sd):
    from dffinthewild_b200.Depth_Estimation_Network import Network
    torch.manual_seed(0)
    net = Network()
    net.load_state_dict(sd, strict=True)
    return net.cuda().train()


def _loss(outs, gt, mask):
    crit = torch.nn.MSELoss()
    return 0.5 * crit(outs[1][mask], gt[mask]) + 0.7 * crit(outs[2][mask], gt[mask]) + 1.0 * crit(outs[3][mask], gt[mask]) \
        + 0.3 * crit(outs[0][mask], gt[mask])


def _state():
    from dffinthewild_b200.Depth_Estimation_Network import Network
    from dffinthewild_b200 import synth
    torch.manual_seed(0)
    return synth.synthetic_state(Network().state_dict(), seed=1)


def test_golden_g3_train_step(built_lib):
    from dffinthewild_b200 import synth
    g = golden("g3_train_synth.npz")
    sd = _state()
    net = _net(sd)
    FS, fd = synth.focal_stack(2, 4, 32, 32, seed=13), synth.focus_dists(2, 4, 32, 32, "defocus")
    gt, mask = synth.gt_and_mask(2, 32, 32, seed=13)
    outs = net(FS.cuda(), fd.cuda())
    for o, n in zip(outs, NAMES):
        ref = g[n]
        assert np.abs(o.detach().cpu().numpy() - ref).max() <= 1e-4 * np.abs(ref).max(), n
    loss = _loss(outs, gt.cuda(), mask.cuda())
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    loss.backward()
    params = dict(net.named_parameters())
    # the 12 parameters the reference never reaches get no gradient here either
    assert sorted(k for k, p in params.items() if p.grad is None) == sorted(str(k) for k in g["grad_none"])
    for k, s, a in zip(g["grad_names"], g["grad_sum"], g["grad_abs"]):
        gr = params[str(k)].grad.double()
        assert abs(float(gr.abs().sum()) - a) <= 2e-3 * a + 1e-12, k
        assert abs(float(gr.sum()) - s) <= 2e-3 * a + 1e-12, k
    worst = 1.0
    for key in g.files:
        if key.startswith("grad:"):
            ref = torch.from_numpy(g[key]).double().flatten()
            got = params[key[5:]].grad.double().cpu().flatten()
            cos = float(torch.dot(ref, got) / (ref.norm() * got.norm() + 1e-300))
            worst = min(worst, cos)
            assert cos >= 0.9999, (key, cos)
    # running statistics after one step (momentum 0.1, unbiased variance) as nn.BatchNorm3d leaves them
    new_sd = net.state_dict()
    for key in g.files:
        if key.startswith("bn:"):
            ref = g[key]
            got = new_sd[key[3:]].cpu().numpy()
            assert np.abs(got - ref).max() <= 1e-4 * max(1e-6, np.abs(ref).max()), key
    k = "DFF_net.dres4.conv0.0.1.num_batches_tracked"
    assert int(new_sd[k]) == int(sd[k]) + 1


def test_gradients_vs_fp64_oracle(built_lib):
    """fp32-mode gradient gate of SURVEY.md §8d: per-tensor cosine >= 0.9999 against the fp64 CPU oracle."""
    from oracle import dff_oracle as O
    from dffinthewild_b200 import synth
    sd = _state()
    net = _net(sd)
    B, S, H, W = 2, 5, 64, 32
    FS, fd = synth.focal_stack(B, S, H, W, seed=21), synth.focus_dists(B, S, H, W, "defocus")
    gt, mask = synth.gt_and_mask(B, H, W, seed=21)
    outs = net(FS.cuda(), fd.cuda())
    _loss(outs, gt.cuda(), mask.cuda()).backward()
    sdo = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
           for k, v in sd.items()}
    o = O.dff_forward(sdo, FS.double(), fd.double(), train=True)
    O.defocus_loss(o, gt.double(), mask).backward()
    for a, b, n in zip(outs, o, NAMES):
        rel = ((a.detach().cpu().double() - b.detach()).abs() / b.detach().abs()).max().item()
        assert rel <= 1e-4, (n, rel)
    worst, worst_k = 1.0, None
    for k, p in net.named_parameters():
        if p.grad is None:
            assert sdo[k].grad is None, k
            continue
        ref, got = sdo[k].grad.flatten(), p.grad.double().cpu().flatten()
        cos = float(torch.dot(ref, got) / (ref.norm() * got.norm() + 1e-300))
        if cos < worst:
            worst, worst_k = cos, k
    assert worst >= 0.9999, (worst_k, worst)


def test_adam_step_runs_on_module_parameters(built_lib):
    """train_code_Defocus.py:67,159-168: zero_grad / backward / Adam.step on the drop-in module's own parameters."""
    from dffinthewild_b200 import synth
    net = _net(_state())
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.9, 0.99))
    FS, fd = synth.focal_stack(1, 3, 32, 32, seed=31).cuda(), synth.focus_dists(1, 3, 32, 32, "defocus").cuda()
    gt, mask = synth.gt_and_mask(1, 32, 32, seed=31)
    before = net.DFF_net.classif3[0].weight.detach().clone()
    losses = []
    for _ in range(3):
        outs = net(FS, fd)
        opt.zero_grad()
        loss = _loss(outs, gt.cuda(), mask.cuda())
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses))
    assert not torch.equal(before, net.DFF_net.classif3[0].weight.detach())
    net.eval()
    with torch.no_grad():
        o = net(FS, fd)          # eval path picks up the updated weights and running statistics
    assert all(torch.isfinite(t).all() for t in o)


def test_bf16_train_step_tensor_cores(built_lib):
    """bf16 training path: tcgen05 forward + tensor-core data gradients, fp32 weight gradients / BatchNorm statistics.
    Gate (per-tensor cosine is not usable in bf16 at these weights): outputs and loss track the fp32 path,
    gradients are finite and globally aligned with the fp32 gradients, Adam steps stay finite."""
    from dffinthewild_b200 import synth
    sd = _state()
    B, S, H, W = 2, 5, 64, 64
    FS, fd = synth.focal_stack(B, S, H, W, seed=41).cuda(), synth.focus_dists(B, S, H, W, "defocus").cuda()
    gt, mask = synth.gt_and_mask(B, H, W, seed=41)
    gt, mask = gt.cuda(), mask.cuda()
    res = {}
    for prec in ("fp32", "bf16"):
        net = _net(sd)
        net.DFF_net.precision = prec
        outs = net(FS, fd)
        loss = _loss(outs, gt, mask)
        loss.backward()
        res[prec] = (loss.item(), [o.detach().float() for o in outs],
                     torch.cat([p.grad.float().flatten() for p in net.parameters() if p.grad is not None]))
    l32, o32, g32 = res["fp32"]
    l16, o16, g16 = res["bf16"]
    assert abs(l16 - l32) <= 3e-2 * abs(l32), (l16, l32)
    for a, b in zip(o16, o32):
        assert ((a - b).abs().mean() / b.abs().mean()).item() <= 5e-2
    assert torch.isfinite(g16).all()
    cos = float(torch.dot(g16.double(), g32.double()) / (g16.double().norm() * g32.double().norm()))
    assert cos >= 0.8, cos
    net = _net(sd)
    net.DFF_net.precision = "bf16"
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.9, 0.99))
    for _ in range(3):
        opt.zero_grad()
        loss = _loss(net(FS, fd), gt, mask)
        loss.backward()
        opt.step()
        assert np.isfinite(loss.item())


def test_golden_g6_train_step_c3_shape(built_lib):
    """Train-mode forward + loss + gradients at BASELINE configs[2]'s stack shape (2 x 5x3x256x256) against the reference's
    committed golden (tests/golden/g6_train_c3.npz): the shapes the training benchmark runs, not only 32x32 toys."""
    from dffinthewild_b200 import synth
    g = golden("g6_train_c3.npz")
    net = _net(_state())
    FS, fd = synth.focal_stack(2, 5, 256, 256, seed=16), synth.focus_dists(2, 5, 256, 256, "defocus")
    gt, mask = synth.gt_and_mask(2, 256, 256, seed=16)
    outs = net(FS.cuda(), fd.cuda())
    for o, n, s in zip(outs, NAMES, g["out_sum"]):
        ref = g[n]
        got = o.detach()[:, ::8, ::8].cpu().numpy()
        assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max(), n
        assert abs(float(o.detach().double().sum()) - s) <= 1e-5 * abs(s), n
    loss = _loss(outs, gt.cuda(), mask.cuda())
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    loss.backward()
    params = dict(net.named_parameters())
    for k, a, l2 in zip(g["grad_names"], g["grad_abs"], g["grad_l2"]):
        gr = params[str(k)].grad.double()
        # (sums of ~650 k cancelling per-pixel terms in fp32: the bias gradients of the first layers move by ~0.5 % with the
        # summation order; the per-tensor direction is gated by the cosine below)
        assert abs(float(gr.abs().sum()) - a) <= 1e-2 * a + 1e-12, k
        assert abs(float(gr.norm()) - l2) <= 1e-2 * l2 + 1e-12, k
    for key in g.files:
        if key.startswith("grad:"):
            ref = torch.from_numpy(g[key]).double().flatten()
            got = params[key[5:]].grad.double().cpu().flatten()
            cos = float(torch.dot(ref, got) / (ref.norm() * got.norm() + 1e-300))
            # (8-element bias gradients that are sums of 655 k cancelling terms: fp32 summation order alone moves their direction
            # by ~1e-4; weight tensors keep the 0.9999 gate)
            assert cos >= (0.9995 if ref.numel() <= 32 else 0.9999), (key, cos)
    new_sd = net.state_dict()
    for key in g.files:
        if key.startswith("bn:"):
            ref = g[key]
            assert np.abs(new_sd[key[3:]].cpu().numpy() - ref).max() <= 1e-4 * max(1e-6, np.abs(ref).max()), key


def test_masked_mse_matches_reference_recipe(built_lib):
    """dff_masked_mse = 0.5*MSE(pred1[mask], gt[mask]) + ... (train_code_Defocus.py:160-165), value and gradients."""
    from dffinthewild_b200 import synth
    from dffinthewild_b200 import train_step as TS
    gt, mask = synth.gt_and_mask(3, 40, 56, seed=51)
    gen = torch.Generator().manual_seed(52)
    preds = [(gt + 0.3 * torch.randn(gt.shape, generator=gen)).cuda().requires_grad_(True) for _ in range(4)]
    ref_preds = [p.detach().clone().double().requires_grad_(True) for p in preds]
    w = (0.3, 0.5, 0.7, 1.0)
    loss = TS.masked_mse_loss(preds, gt.cuda(), mask.cuda(), w)
    (loss * 1.7).backward()
    crit = torch.nn.MSELoss()
    ref = sum(wk * crit(p[mask.cuda()], gt.cuda().double()[mask.cuda()]) for wk, p in zip(w, ref_preds))
    (ref * 1.7).backward()
    assert abs(loss.item() - ref.item()) <= 1e-6 * abs(ref.item())
    for p, r in zip(preds, ref_preds):
        assert (p.grad.double() - r.grad).abs().max().item() <= 1e-6 * r.grad.abs().max().item()
    # no valid pixel: zero loss and zero gradients instead of 0/0
    z = TS.masked_mse_loss([p.detach() for p in preds], gt.cuda(), torch.zeros_like(mask).cuda(), w)
    assert z.item() == 0.0


def test_adam_flat_is_bit_compatible_with_torch(built_lib):
    """dff_adam_flat against torch.optim.Adam (the CUDA default, foreach implementation) with the reference's hyper-parameters
    (train_code_Defocus.py:67): parameters and both moments bit-identical after 3 steps."""
    import ctypes
    from dffinthewild_b200 import runtime as rt
    from dffinthewild_b200 import train_step as TS
    l = rt.lib()
    TS._declare(l)
    gen = torch.Generator().manual_seed(61)
    n = 100003
    p0 = torch.randn(n, generator=gen).cuda()
    grads = [(torch.randn(n, generator=gen) * (10.0 ** torch.randint(-6, 2, (n,), generator=gen).float())).cuda() for _ in range(3)]
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.99), foreach=True)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for t, g in enumerate(grads, 1):
        ref.grad = g.clone()
        opt.step()
        rt.check(l.dff_adam_flat(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.99, 1e-8, t, None, 0,
                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    st = opt.state[ref]
    assert torch.equal(m, st["exp_avg"]) and torch.equal(v, st["exp_avg_sq"])
    assert torch.equal(p, ref.detach())


# Tolerances: the first steps of Adam move every weight by ~lr * sign(g), so fp32 summation-order noise in near-zero gradients turns
# into O(lr) = 1e-3 weight differences after one step: step 0 agrees to 1e-4, steps 1-2 to a few 1e-3 (fp32) / 3 % (bf16).
@pytest.mark.parametrize("precision,rtol", [("fp32", 5e-3), ("bf16", 3e-2)])
def test_golden_g7_training_loop_trajectory(built_lib, precision, rtol):
    """Three optimizer steps through TrainStep (fused masked-MSE, flat Adam) against the reference's own loop
    (tests/golden/g7_train_3steps.npz): the loss trajectory — SURVEY.md §8(d)'s bf16 training gate — and, in fp32, the
    parameters and running statistics after the third step."""
    from dffinthewild_b200 import synth
    from dffinthewild_b200 import train_step as TS
    g = golden("g7_train_3steps.npz")
    net = _net(_state())
    net.DFF_net.precision = precision
    stepper = TS.TrainStep(net, lr=1e-3, betas=(0.9, 0.99), weights=(0.3, 0.5, 0.7, 1.0))
    for step in range(3):
        FS, fd = synth.focal_stack(2, 4, 32, 32, seed=20 + step), synth.focus_dists(2, 4, 32, 32, "defocus")
        gt, mask = synth.gt_and_mask(2, 32, 32, seed=20 + step)
        info = stepper.step(FS.cuda(), fd.cuda(), gt.cuda(), mask.cuda())
        tol = 1e-4 if (step == 0 and precision == "fp32") else rtol
        assert abs(float(info["loss"]) - g["losses"][step]) <= tol * g["losses"][step], (step, float(info["loss"]), g["losses"][step])
    if precision == "fp32":
        sd = net.state_dict()
        for k in ("DFF_net.classif3.0.weight", "DFF_net.dres4.conv6.1.weight", "DFF_net.dres4.conv6.1.running_var"):
            ref = g["w:" + k]
            assert np.abs(sd[k].cpu().numpy() - ref).max() <= 1e-2 * np.abs(ref).max(), k
        # the module still answers in eval mode with the stepped weights (flat-buffer views + invalidated weight pack)
        net.eval()
        with torch.no_grad():
            o = net(FS.cuda(), fd.cuda())
        assert all(torch.isfinite(t).all() for t in o)


WGRAD_CASES = [
    # name, C0, C1, Cout, k, stride, dil, transposed, S, H, W
    ("c3_16_8", 16, 0, 8, (3, 3, 3), 1, 1, False, 3, 40, 72),
    ("c3_8+8_8_two_sources", 8, 8, 8, (3, 3, 3), 1, 1, False, 3, 32, 64),
    ("srd_1x3x3_8_8_tap_pairs", 8, 0, 8, (1, 3, 3), 1, 1, False, 2, 32, 40),
    ("fm_9x9_dil2_first_layer", 8, 0, 8, (1, 9, 9), 1, 2, False, 2, 32, 64),
    ("s2_8_16", 8, 0, 16, (3, 3, 3), 2, 1, False, 3, 64, 96),
    ("s2_32_64", 32, 0, 64, (3, 3, 3), 2, 1, False, 2, 24, 40),
    ("c3_64_64", 64, 0, 64, (3, 3, 3), 1, 1, False, 3, 12, 20),
    ("c3_128+64_128", 128, 64, 128, (3, 3, 3), 1, 1, False, 2, 6, 10),
    ("up_32_16", 32, 0, 16, (3, 3, 3), 2, 1, True, 3, 16, 24),
    ("up_16_8", 16, 0, 8, (3, 3, 3), 2, 1, True, 2, 32, 48),
    ("att_3x1x1_16", 16, 0, 16, (3, 1, 1), 1, 1, False, 5, 16, 48),
    ("cls_1x1x1_8_1", 8, 0, 1, (1, 1, 1), 1, 1, False, 2, 32, 64),
]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=[c[0] for c in WGRAD_CASES])
def test_wgrad_tensor_core_kernel(built_lib, case):
    """bf16 weight gradient (C-ABI dff_conv3d_wgrad, warp-level tensor-core kernel) against torch's fp64 convolution weight
    gradient on the same bf16-rounded operands: exact products, fp32 accumulation -> 2e-5 of the largest entry (the only error is
    fp32 summation order over B*S*OH*OW terms); every layer class of the network incl. strided, transposed, two-source, dilated."""
    import ctypes
    import torch.nn.functional as F
    from dffinthewild_b200 import runtime as rt
    from dffinthewild_b200 import train as tr
    name, c0, c1, cout, k, stride, dil, transposed, S, H, W = case
    l = tr._lib()
    B = 2
    g = torch.Generator().manual_seed(71)
    cin = c0 + c1
    x = (torch.rand(B, cin, S, H, W, generator=g) * 2 - 1).bfloat16().float()
    OH, OW = (2 * H, 2 * W) if transposed else (H // stride, W // stride)
    cos = max(8, (cout + 7) // 8 * 8)
    dy = (torch.rand(B, cout, S, OH, OW, generator=g) * 2 - 1).bfloat16().float()
    wshape = (cin, cout) + k if transposed else (cout, cin) + k
    w = torch.zeros(wshape, dtype=torch.float64, requires_grad=True)
    if transposed:
        y = F.conv_transpose3d(x.double(), w, None, stride=(1, 2, 2), padding=1, output_padding=(0, 1, 1))
    else:
        pad = ((k[0] - 1) // 2, dil * (k[1] - 1) // 2, dil * (k[2] - 1) // 2)
        y = F.conv3d(x.double(), w, None, (1, stride, stride), pad, (1, dil, dil))
    (y * dy.double()).sum().backward()
    ref = w.grad
    x0 = rt.to_channels_last(x[:, :c0].cuda().contiguous(), c0, True)
    x1 = rt.to_channels_last(x[:, c0:].cuda().contiguous(), c1, True) if c1 else None
    dyc = rt.to_channels_last(dy.cuda(), cos, True)
    dw = torch.full(wshape, float("nan"), dtype=torch.float32, device="cuda")
    rt.check(l.dff_conv3d_wgrad(x0.data_ptr(), c0, x1.data_ptr() if c1 else None, c1, B, S, H, W, dyc.data_ptr(), cos, cin, cout, k[0], k[1],
                                k[2], 2 if transposed else stride, dil, 1 if transposed else 0, dw.data_ptr(), rt.BF16, 0,
                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    err = (dw.cpu().double() - ref).abs().max().item()
    assert err <= 2e-5 * ref.abs().max().item(), (name, err, ref.abs().max().item())
