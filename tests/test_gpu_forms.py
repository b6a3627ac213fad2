"""Kernel-level bf16 parity of every FORM `dff_forward` can pick for a layer (C-ABI `dff_conv3d_ex`, plan 4 = the forward's own
kernel selection): x-folded small-Cout convolutions (G = 2 / 4), x-folded transposed convolutions, the row-folded pair-packed
first layer, the second output (`x + out`, reference train_codes/Depth_Estimation_Network.py:104,110) and the fused 1x1x1
classifiers (reference :53-57,105,111,116).  Reference = torch fp64 on the same bf16-rounded operands: what is left is fp32
accumulation order and the bf16 rounding of the stored result, so the gate is 6e-3 * max|ref| as in test_conv_tcgen05_bf16."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 6e-3


@pytest.fixture(scope="module")
def rt(built_lib):
    from dffinthewild_b200 import runtime
    assert torch.cuda.is_available()
    return runtime


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(*shape, generator=g) * 2 - 1) * scale).bfloat16().float()


def _ref_conv(x, w, stride, dil, transposed):
    x, w = x.double(), w.double()
    if transposed:
        return F.conv_transpose3d(x, w, None, stride=(1, 2, 2), padding=1, output_padding=(0, 1, 1))
    k = w.shape[2:]
    pad = ((k[0] - 1) // 2, dil * (k[1] - 1) // 2, dil * (k[2] - 1) // 2)
    return F.conv3d(x, w, None, (1, stride, stride), pad, (1, dil, dil))


def _ss(cout):
    g = torch.Generator().manual_seed(3)
    return torch.rand(cout, generator=g) * 0.8 + 0.6, (torch.rand(cout, generator=g) - 0.5) * 0.6


def _bc(v):
    return v.double().view(1, -1, 1, 1, 1)


# (name, C0, C1, Cout, k, dil, S, H, W): the network's layers that run x-folded — G = 4: 8 -> 8; G = 2: 16 -> 8 and Cin >= 32 -> 16
XFOLD = [
    ("srd8_1x3x3_G4", 8, 0, 8, (1, 3, 3), 1, 3, 32, 64),          # FM_measure...Focus_Measure.conv.0 / conv.2
    ("c3_8_8_G4", 8, 0, 8, (3, 3, 3), 1, 4, 16, 96),
    ("dres4_conv0_8+8_8_G2", 8, 8, 8, (3, 3, 3), 1, 3, 32, 64),   # hourglass(8).conv0 on cat[out2, V1]
    ("dres3_conv0_16+16_16_G2", 16, 16, 16, (3, 3, 3), 1, 3, 16, 48),
    ("c3_32_16_G2", 32, 0, 16, (3, 3, 3), 1, 2, 24, 32),
    ("c3_8_8_G4_1slice", 8, 0, 8, (3, 3, 3), 1, 1, 32, 32),
]


@pytest.mark.parametrize("case", XFOLD, ids=[c[0] for c in XFOLD])
def test_xfolded_conv(rt, case):
    name, c0, c1, cout, k, dil, S, H, W = case
    B = 2
    x0 = _rand(B, c0, S, H, W, seed=1)
    x1 = _rand(B, c1, S, H, W, seed=2) if c1 else None
    cin = c0 + c1
    w = _rand(cout, cin, *k, seed=3, scale=(2.0 / (cin * k[0] * k[1] * k[2])) ** 0.5 * 1.7)
    ref = _ref_conv(torch.cat([x0, x1], 1) if c1 else x0, w, 1, dil, False)
    out = rt.conv3d_forward_plan(x0.cuda(), w.cuda(), x2=x1.cuda() if c1 else None)["out"].cpu().double()
    assert (out - ref).abs().max().item() <= TOL * ref.abs().max().item(), name
    # the epilogues these layers run with: BN + ReLU (conv.0, conv0), BN + residual + ReLU (conv.2)
    scale, shift = _ss(cout)
    res = _rand(*ref.shape, seed=5)
    full = F.relu(ref * _bc(scale) + _bc(shift) + res.double())
    got = rt.conv3d_forward_plan(x0.cuda(), w.cuda(), x2=x1.cuda() if c1 else None, scale=scale.cuda(), shift=shift.cuda(),
                                 res_pre=res.cuda(), relu=True)["out"].cpu().double()
    assert (got - full).abs().max().item() <= TOL * full.abs().max().item(), name


def test_xfold_agrees_with_plain_slab_form(rt):
    """The folded and the plain form of the same layer are the same class of numbers (exact products, fp32 accumulate, one bf16
    rounding): they may differ by accumulation order only."""
    x, w = _rand(1, 8, 2, 16, 64, seed=7), _rand(8, 8, 1, 3, 3, seed=8, scale=0.3)
    a = rt.conv3d_forward_plan(x.cuda(), w.cuda())["out"].cpu()
    b = rt.conv3d(x.cuda(), w.cuda(), bf16=True, tensor_cores=3).cpu()     # plain slab form
    assert (a - b).abs().max().item() <= 2.0 ** -7 * b.abs().max().item()


# transposed convolutions with Cout <= 32 run x-folded (two row phases, two column phases in one GEMM row)
DECONV = [
    ("deconv_1_64_32", 64, 32, 2, 12, 16),
    ("dres3_conv5_32_32", 32, 32, 3, 16, 24),
    ("deconv_2_32_16", 32, 16, 3, 16, 32),
    ("dres4_conv5_16_16", 16, 16, 2, 32, 32),
    ("deconv_3_16_8", 16, 8, 3, 32, 48),
]


@pytest.mark.parametrize("case", DECONV, ids=[c[0] for c in DECONV])
def test_xfolded_transposed_conv(rt, case):
    name, cin, cout, S, H, W = case
    B = 2
    x = _rand(B, cin, S, H, W, seed=11)
    w = _rand(cin, cout, 3, 3, 3, seed=12, scale=(2.0 / (cin * 27 / 4)) ** 0.5 * 1.5)
    ref = _ref_conv(x, w, 2, 1, True)
    out = rt.conv3d_forward_plan(x.cuda(), w.cuda(), 2, 1, True)["out"].cpu().double()
    assert (out - ref).abs().max().item() <= TOL * ref.abs().max().item(), name
    scale, shift = _ss(cout)
    res = _rand(*ref.shape, seed=13)
    full = F.relu(ref * _bc(scale) + _bc(shift) + res.double())          # conv5: deconv -> BN -> + presqu -> ReLU (reference :314-315)
    got = rt.conv3d_forward_plan(x.cuda(), w.cuda(), 2, 1, True, scale=scale.cuda(), shift=shift.cuda(), res_pre=res.cuda(),
                                 relu=True)["out"].cpu().double()
    assert (got - full).abs().max().item() <= TOL * full.abs().max().item(), name


CONV6 = [("dres2_conv6_64_32", 64, 32, 2, 12, 16), ("dres3_conv6_32_16", 32, 16, 3, 16, 24), ("dres4_conv6_16_8", 16, 8, 2, 32, 32)]


@pytest.mark.parametrize("case", CONV6, ids=[c[0] for c in CONV6])
def test_conv6_second_output_and_fused_classifier(rt, case):
    """hourglass.conv6 as the forward runs it: out = BN(deconv(x)); out_in = skip + out (second output); cost = classif(out_in)."""
    name, cin, cout, S, H, W = case
    B = 2
    x = _rand(B, cin, S, H, W, seed=21)
    w = _rand(cin, cout, 3, 3, 3, seed=22, scale=(2.0 / (cin * 27 / 4)) ** 0.5 * 1.5)
    scale, shift = _ss(cout)
    ref = _ref_conv(x, w, 2, 1, True) * _bc(scale) + _bc(shift)
    skip = _rand(*ref.shape, seed=23, scale=2.0)
    pw = _rand(cout, seed=24, scale=0.7)
    r = rt.conv3d_forward_plan(x.cuda(), w.cuda(), 2, 1, True, scale=scale.cuda(), shift=shift.cuda(), aux_add=skip.cuda(),
                               proj_w=pw.cuda(), proj_on_aux=True)
    out, aux, proj = r["out"].cpu().double(), r["aux"].cpu().double(), r["proj"].cpu().double()
    assert (out - ref).abs().max().item() <= TOL * ref.abs().max().item(), name
    aux_ref = ref + skip.double()
    assert (aux - aux_ref).abs().max().item() <= TOL * aux_ref.abs().max().item(), name
    # the classifier sees exactly the bf16 values the next layer will read: tight against the kernel's own second output
    proj_ref = (aux * _bc(pw)).sum(1)
    assert (proj - proj_ref).abs().max().item() <= 1e-5 * (aux.abs() * _bc(pw).abs()).sum(1).max().item(), name
    # last stage (reference :115-116): only the cost is needed — out2 + out is added after BN (res_post) and never stored
    r2 = rt.conv3d_forward_plan(x.cuda(), w.cuda(), 2, 1, True, scale=scale.cuda(), shift=shift.cuda(), res_post=skip.cuda(),
                                proj_w=pw.cuda(), proj_on_aux=False, skip_out=True)
    assert "out" not in r2
    bound = 2.0 ** -8 * (aux_ref.abs() * _bc(pw).abs()).sum(1) + 1e-4      # each term rounded to bf16 before the dot product
    assert bool(((r2["proj"].cpu().double() - (aux_ref * _bc(pw)).sum(1)).abs() <= bound).all()), name


@pytest.mark.parametrize("S,H,W", [(2, 32, 64), (3, 40, 96), (1, 64, 32)])
def test_rowfolded_first_layer(rt, S, H, W):
    """FM_measure's 1x9x9 dilation-2 conv on the pair-packed input, four output rows per GEMM row (reference :144-148)."""
    B = 2
    x = _rand(B, 3, S, H, W, seed=31)
    x[:, :, :, H - 3:, :] = -1.0      # the dataloader's -1 padding
    w = _rand(8, 3, 1, 9, 9, seed=32, scale=(2.0 / 243) ** 0.5 * 1.7)
    scale, shift = _ss(8)
    ref = F.relu(_ref_conv(x, w, 1, 2, False) * _bc(scale) + _bc(shift))
    out = rt.conv3d_forward_plan(x.cuda(), w.cuda(), 1, 2, False, scale=scale.cuda(), shift=shift.cuda(), relu=True)["out"].cpu().double()
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= TOL * ref.abs().max().item()


def test_streamed_weights_and_per_tap_leftovers(rt):
    """The layers whose weights do not fit next to the plane ring (64 -> 64, stride-2 32 -> 64) and the >= 128-channel ones."""
    for cin, cout, stride, S, H, W in ((64, 64, 1, 3, 16, 24), (32, 64, 2, 2, 32, 48), (128, 64, 1, 2, 16, 16), (192, 128, 1, 2, 8, 16)):
        x = _rand(1, cin, S, H, W, seed=41)
        w = _rand(cout, cin, 3, 3, 3, seed=42, scale=(2.0 / (cin * 27)) ** 0.5 * 1.7)
        ref = F.relu(_ref_conv(x, w, stride, 1, False))
        out = rt.conv3d_forward_plan(x.cuda(), w.cuda(), stride, 1, False, relu=True)["out"].cpu().double()
        assert (out - ref).abs().max().item() <= TOL * ref.abs().max().item(), (cin, cout, stride)


# (name, C0, C1, Cout, k, stride, S, H, W): sources of 16 / 32 / 64n channels run the wide-row plane layout (pixel-major planes,
# 32B / 64B / 128B-swizzled TMA boxes and A descriptors): every swizzle width, concat sources, the four stride-2 views, 1x3x3 and
# focal-merged 3x3x3 schedules, resident and streamed weights, sizes whose tiles are ragged in x and y (halo / out-of-bounds fill)
WIDE_ROW = [
    ("wr32B_16_16_1x3x3", 16, 0, 16, (1, 3, 3), 1, 3, 24, 40),
    ("wr32B_16_32_s2", 16, 0, 32, (3, 3, 3), 2, 3, 32, 48),
    ("wr32B_16+16_16_ragged", 16, 16, 16, (3, 3, 3), 1, 2, 20, 44),
    ("wr64B_32_32_1x3x3", 32, 0, 32, (1, 3, 3), 1, 2, 16, 24),
    ("wr64B_32+32_32", 32, 32, 32, (3, 3, 3), 1, 3, 16, 24),
    ("wr64B_32_64_s2", 32, 0, 64, (3, 3, 3), 2, 2, 16, 32),
    ("wr128B_64_64_1x3x3", 64, 0, 64, (1, 3, 3), 1, 2, 8, 24),
    ("wr128B_64+64_64", 64, 64, 64, (3, 3, 3), 1, 2, 8, 16),
    ("wr128B_128+64_128", 128, 64, 128, (3, 3, 3), 1, 2, 8, 8),
    ("wr32B_16_8_1x1x1", 16, 0, 8, (1, 1, 1), 1, 2, 16, 32),
    # stride-2 layers with 8 / 16-channel sources run x-paired: the source read as (.., W/2, 2C), two row-parity views
    ("xpair_8_16_s2", 8, 0, 16, (3, 3, 3), 2, 3, 32, 48),
    ("xpair_8_16_s2_ragged", 8, 0, 16, (3, 3, 3), 2, 2, 20, 44),
    ("xpair_16_16_s2", 16, 0, 16, (3, 3, 3), 2, 2, 24, 40),
    ("xpair_8_8_1x3x3_s2", 8, 0, 8, (1, 3, 3), 2, 2, 16, 32),
    # x-grouped forms (widths that are multiples of 8 G output pixels): the source read as groups of four pixels, banded group-tap weights,
    # all-zero K steps not issued
    ("xgroup_8_16_s2", 8, 0, 16, (3, 3, 3), 2, 3, 32, 64),
    ("xgroup_8_8_s2", 8, 0, 8, (3, 3, 3), 2, 2, 16, 32),
    ("xgroup_8_16_s2_1x3x3", 8, 0, 16, (1, 3, 3), 2, 2, 20, 96),
    ("xgroup_8_8_1x3x3_G4", 8, 0, 8, (1, 3, 3), 1, 3, 24, 64),
    ("xgroup_8_8_1x3x3_G4_wide", 8, 0, 8, (1, 3, 3), 1, 2, 12, 160),
    ("xgroup_16_16_1x3x3_G2", 16, 0, 16, (1, 3, 3), 1, 3, 24, 64),
    ("xgroup_16_16_1x3x3_G2_wide", 16, 0, 16, (1, 3, 3), 1, 2, 20, 112),
]


@pytest.mark.parametrize("case", WIDE_ROW, ids=[c[0] for c in WIDE_ROW])
def test_wide_row_plane_layout(rt, case):
    """bf16 convolution with BN scale/shift, residual before the ReLU, on 16/32/64n-channel sources (wide-row planes) against torch
    fp64 on the same bf16-rounded operands."""
    name, c0, c1, cout, k, stride, S, H, W = case
    cin = c0 + c1
    x = _rand(2, cin, S, H, W, seed=51)
    w = _rand(cout, cin, *k, seed=52, scale=(2.0 / (cin * k[0] * k[1] * k[2])) ** 0.5 * 1.7)
    scale, shift = _ss(cout)
    res = _rand(2, cout, S, H // stride, W // stride, seed=53)
    ref = F.relu(_ref_conv(x, w, stride, 1, False) * _bc(scale) + _bc(shift) + res.double())
    out = rt.conv3d_forward_plan(x[:, :c0].contiguous().cuda(), w.cuda(), stride, 1, False, scale=scale.cuda(), shift=shift.cuda(),
                                 res_pre=res.cuda(), relu=True, x2=x[:, c0:].contiguous().cuda() if c1 else None)["out"].cpu().double()
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= TOL * ref.abs().max().item(), name
