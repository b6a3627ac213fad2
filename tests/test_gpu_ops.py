"""Kernel-level parity on the GPU, called through the C-ABI: every conv flavour of the network against
torch's fp64 CPU convolution (the primitive the reference calls), the depth head against the oracle, the FOV warp
against the reference's golden outputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rt(built_lib):
    from dffinthewild_b200 import runtime
    assert torch.cuda.is_available()
    return runtime


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g) * 2 - 1) * scale


CASES = [
    # name, Cin, Cout, k, stride, dil, transposed, S, H, W
    ("fm_9x9_dil2", 3, 8, (1, 9, 9), 1, 2, False, 2, 40, 72),
    ("srd_1x3x3_c8", 8, 8, (1, 3, 3), 1, 1, False, 3, 32, 40),
    ("att_3x1x1_c16", 16, 16, (3, 1, 1), 1, 1, False, 5, 16, 48),
    ("att_1x1x1_c32", 32, 32, (1, 1, 1), 1, 1, False, 2, 16, 16),
    ("c3_32_32", 32, 32, (3, 3, 3), 1, 1, False, 4, 24, 40),
    ("c3_64_64", 64, 64, (3, 3, 3), 1, 1, False, 3, 12, 20),
    ("c3_128_128", 128, 128, (3, 3, 3), 1, 1, False, 2, 6, 10),
    ("c3_16_8", 16, 8, (3, 3, 3), 1, 1, False, 3, 64, 64),
    ("s2_8_16", 8, 16, (3, 3, 3), 2, 1, False, 3, 64, 96),
    ("s2_32_64", 32, 64, (3, 3, 3), 2, 1, False, 2, 24, 40),
    ("s2_64_128", 64, 128, (3, 3, 3), 2, 1, False, 1, 12, 20),
    ("up_64_32", 64, 32, (3, 3, 3), 2, 1, True, 3, 12, 20),
    ("up_16_8", 16, 8, (3, 3, 3), 2, 1, True, 2, 32, 48),
    ("up_128_64", 128, 64, (3, 3, 3), 2, 1, True, 2, 3, 5),
    ("conf_32_1", 32, 1, (3, 3, 3), 1, 1, False, 3, 12, 20),
    ("cls_8_1", 8, 1, (1, 1, 1), 1, 1, False, 2, 32, 64),
]


def _ref_conv(x, w, stride, dil, transposed):
    x, w = x.double(), w.double()
    if transposed:
        return F.conv_transpose3d(x, w, None, stride=(1, 2, 2), padding=1, output_padding=(0, 1, 1))
    k = w.shape[2:]
    pad = ((k[0] - 1) // 2, dil * (k[1] - 1) // 2, dil * (k[2] - 1) // 2)
    return F.conv3d(x, w, None, (1, stride, stride), pad, (1, dil, dil))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_fp32_matches_torch_fp64(rt, case):
    name, cin, cout, k, stride, dil, tr, S, H, W = case
    B = 2
    x = _rand(B, cin, S, H, W, seed=1)
    wshape = (cin, cout) + k if tr else (cout, cin) + k
    w = _rand(*wshape, seed=2, scale=(2.0 / (cin * k[0] * k[1] * k[2])) ** 0.5 * 1.7)
    ref = _ref_conv(x, w, stride, dil, tr)
    scale = _rand(cout, seed=3) * 0.4 + 1.0
    shift = _rand(cout, seed=4) * 0.3
    res1 = _rand(*ref.shape, seed=5)
    res2 = _rand(*ref.shape, seed=6)
    full = F.relu(ref * scale.double().view(1, -1, 1, 1, 1) + shift.double().view(1, -1, 1, 1, 1) + res1.double()) + res2.double()
    out_plain = rt.conv3d(x.cuda(), w.cuda(), stride, dil, tr).cpu().double()
    if cout % 4:   # C -> 1 projections (confidence.2, classif*): no residual operands in the network either
        res1, res2 = None, None
        full = F.relu(ref * scale.double().view(1, -1, 1, 1, 1) + shift.double().view(1, -1, 1, 1, 1))
    out_full = rt.conv3d(x.cuda(), w.cuda(), stride, dil, tr, scale=scale.cuda(), shift=shift.cuda(),
                         res_pre=res1.cuda() if res1 is not None else None,
                         res_post=res2.cuda() if res2 is not None else None, relu=True).cpu().double()
    K = cin * k[0] * k[1] * k[2]
    tol = 2e-7 * K ** 0.5 * max(1.0, ref.abs().max().item())   # sequential fp32 accumulation over K terms
    assert (out_plain - ref).abs().max().item() <= tol
    assert (out_full - full).abs().max().item() <= 2 * tol


TC_CASES = [c for c in CASES if c[2] % 8 == 0]


@pytest.mark.parametrize("kernel", [1, 2], ids=["slab", "per_tap_tma"])
@pytest.mark.parametrize("case", TC_CASES, ids=[c[0] for c in TC_CASES])
def test_conv_tcgen05_bf16(rt, case, kernel):
    """tcgen05/TMA implicit-GEMM path: exact products of the bf16-rounded operands, fp32 accumulate, bf16 store."""
    name, cin, cout, k, stride, dil, tr, S, H, W = case
    B = 2
    x = _rand(B, cin, S, H, W, seed=1).bfloat16().float()
    wshape = (cin, cout) + k if tr else (cout, cin) + k
    w = (_rand(*wshape, seed=2, scale=(2.0 / (cin * k[0] * k[1] * k[2])) ** 0.5 * 1.7)).bfloat16().float()
    ref = _ref_conv(x, w, stride, dil, tr)
    out = rt.conv3d(x.cuda(), w.cuda(), stride, dil, tr, bf16=True, tensor_cores=kernel).cpu().double()
    scale_ = ref.abs().max().item()
    assert (out - ref).abs().max().item() <= 6e-3 * scale_, name
    scale = _rand(cout, seed=3) * 0.4 + 1.0
    shift = _rand(cout, seed=4) * 0.3
    res1 = _rand(*ref.shape, seed=5).bfloat16().float()
    res2 = _rand(*ref.shape, seed=6).bfloat16().float()
    full = F.relu(ref * scale.double().view(1, -1, 1, 1, 1) + shift.double().view(1, -1, 1, 1, 1) + res1.double()) + res2.double()
    out_full = rt.conv3d(x.cuda(), w.cuda(), stride, dil, tr, scale=scale.cuda(), shift=shift.cuda(), res_pre=res1.cuda(),
                         res_post=res2.cuda(), relu=True, bf16=True, tensor_cores=kernel).cpu().double()
    assert (out_full - full).abs().max().item() <= 6e-3 * full.abs().max().item(), name


ROW_CASES = [
    # name, C0, C1, Cout, k, S, H, W   (stride 1; W >= 256 selects the row kernel: A operand in tensor memory)
    ("row_c3_16_16", 16, 0, 16, (3, 3, 3), 4, 9, 288),
    ("row_c3_8+8_8", 8, 8, 8, (3, 3, 3), 3, 8, 320),
    ("row_c3_16+16_16", 16, 16, 16, (3, 3, 3), 5, 7, 256),
    ("row_c3_32_32", 32, 0, 32, (3, 3, 3), 3, 8, 288),
    ("row_c3_32+32_32", 32, 32, 32, (3, 3, 3), 2, 6, 264),
    ("row_1x3x3_8_8", 8, 0, 8, (1, 3, 3), 3, 12, 576),
    ("row_1x3x3_16_16", 16, 0, 16, (1, 3, 3), 2, 9, 288),
    ("row_1x3x3_32_32", 32, 0, 32, (1, 3, 3), 2, 5, 256),
    ("row_c3_16_16_1slice", 16, 0, 16, (3, 3, 3), 1, 4, 256),
]


@pytest.fixture()
def row_kernel_on():
    """The row kernel is opt-in (DFF_B200_ROW=1, read once per process): these tests run it in a child interpreter."""
    import os
    import subprocess
    import sys

    def run(case_name):
        env = dict(os.environ, DFF_B200_ROW="1", DFF_ROW_CHILD=case_name)
        r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", __file__, "-k", "row_kernel_child", "-m", "gpu"], env=env,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    return run


@pytest.mark.parametrize("case", ROW_CASES, ids=[c[0] for c in ROW_CASES])
def test_conv_row_kernel_tmem_operand(row_kernel_on, case):
    row_kernel_on(case[0])


def test_row_kernel_child(rt):
    """Body of the row-kernel parity test; only runs inside the child process started by the fixture above."""
    import os
    name = os.environ.get("DFF_ROW_CHILD")
    if not name or os.environ.get("DFF_B200_ROW") != "1":
        pytest.skip("row-kernel child only")
    _row_case(rt, next(c for c in ROW_CASES if c[0] == name))


def _row_case(rt, case):
    """Row kernel (tcgen05.mma TS form, input-stationary over 9 accumulators): must agree with the slab kernel's contract —
    exact products of bf16 operands, fp32 accumulate, fused BN/residual/ReLU epilogue, bf16 store."""
    name, c0, c1, cout, k, S, H, W = case
    B = 2
    x0 = _rand(B, c0, S, H, W, seed=1).bfloat16().float()
    x1 = _rand(B, c1, S, H, W, seed=2).bfloat16().float() if c1 else None
    cin = c0 + c1
    w = _rand(cout, cin, *k, seed=3, scale=(2.0 / (cin * k[0] * k[1] * k[2])) ** 0.5 * 1.7).bfloat16().float()
    xin = torch.cat([x0, x1], 1) if c1 else x0
    ref = _ref_conv(xin, w, 1, 1, False)
    out = rt.conv3d(x0.cuda(), w.cuda(), x2=x1.cuda() if c1 else None, bf16=True, tensor_cores=1).cpu().double()
    assert (out - ref).abs().max().item() <= 6e-3 * ref.abs().max().item(), name
    scale = _rand(cout, seed=4) * 0.4 + 1.0
    shift = _rand(cout, seed=5) * 0.3
    res1 = _rand(*ref.shape, seed=6).bfloat16().float()
    res2 = _rand(*ref.shape, seed=7).bfloat16().float()
    full = F.relu(ref * scale.double().view(1, -1, 1, 1, 1) + shift.double().view(1, -1, 1, 1, 1) + res1.double()) + res2.double()
    out_full = rt.conv3d(x0.cuda(), w.cuda(), x2=x1.cuda() if c1 else None, scale=scale.cuda(), shift=shift.cuda(),
                         res_pre=res1.cuda(), res_post=res2.cuda(), relu=True, bf16=True, tensor_cores=1).cpu().double()
    assert (out_full - full).abs().max().item() <= 6e-3 * full.abs().max().item(), name
    # and bit-for-bit the same result class as the slab kernel (same operands, same accumulation width)
    slab = rt.conv3d(x0.cuda(), w.cuda(), x2=x1.cuda() if c1 else None, bf16=True, tensor_cores=3).cpu().double()
    assert (out - slab).abs().max().item() <= 1.6e-2 * ref.abs().max().item(), name


@pytest.mark.parametrize("kernel", [1, 2], ids=["slab", "per_tap_tma"])
def test_conv_tcgen05_two_sources(rt, kernel):
    for c0, c1, cout in ((16, 16, 16), (8, 8, 8), (32, 32, 32), (64, 64, 64), (128, 64, 128)):
        x0, x1 = _rand(1, c0, 3, 12, 36, seed=7).bfloat16().float(), _rand(1, c1, 3, 12, 36, seed=8).bfloat16().float()
        w = _rand(cout, c0 + c1, 3, 3, 3, seed=9, scale=0.05).bfloat16().float()
        ref = _ref_conv(torch.cat([x0, x1], 1), w, 1, 1, False)
        out = rt.conv3d(x0.cuda(), w.cuda(), x2=x1.cuda(), bf16=True, tensor_cores=kernel).cpu().double()
        assert (out - ref).abs().max().item() <= 6e-3 * ref.abs().max().item(), (c0, c1, cout)


@pytest.mark.parametrize("kernel", [1, 2], ids=["slab", "per_tap_tma"])
def test_conv_tcgen05_many_tiles_persistent(rt, kernel):
    """More tiles than SMs: every CTA walks several tiles through both TMEM accumulator buffers."""
    x = _rand(2, 16, 6, 96, 160, seed=12).bfloat16().float()
    w = _rand(16, 16, 3, 3, 3, seed=13, scale=0.08).bfloat16().float()
    ref = _ref_conv(x, w, 1, 1, False)
    out = rt.conv3d(x.cuda(), w.cuda(), bf16=True, tensor_cores=kernel).cpu().double()
    assert (out - ref).abs().max().item() <= 6e-3 * ref.abs().max().item()


def test_conv_two_sources_is_channel_concat(rt):
    """hourglass conv0 reads torch.cat([decoder, encoder_skip], 1) without materialising it (reference :103,109,114)."""
    for c0, c1, cout in ((16, 16, 16), (8, 8, 8), (64, 64, 64), (128, 64, 128)):
        x0, x1 = _rand(1, c0, 3, 12, 36, seed=7), _rand(1, c1, 3, 12, 36, seed=8)
        w = _rand(cout, c0 + c1, 3, 3, 3, seed=9, scale=0.05)
        ref = _ref_conv(torch.cat([x0, x1], 1), w, 1, 1, False)
        out = rt.conv3d(x0.cuda(), w.cuda(), x2=x1.cuda()).cpu().double()
        assert (out - ref).abs().max().item() <= 2e-7 * (27 * (c0 + c1)) ** 0.5 * max(1.0, ref.abs().max().item())


def test_conv_bf16_storage(rt):
    """bf16 activations in HBM, fp32 accumulate: error bounded by the bf16 rounding of inputs/outputs."""
    x, w = _rand(1, 32, 3, 16, 40, seed=10), _rand(32, 32, 3, 3, 3, seed=11, scale=0.05)
    ref = _ref_conv(x.bfloat16().float(), w, 1, 1, False)
    out = rt.conv3d(x.cuda(), w.cuda(), bf16=True).cpu().double()
    assert (out - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()


@pytest.mark.parametrize("C,S,H,W", [(8, 10, 32, 40), (8, 1, 16, 16), (8, 2, 16, 24), (16, 5, 16, 48), (16, 3, 32, 32), (32, 4, 16, 24),
                                     (32, 10, 32, 32)])
def test_srd_attention_one_pass(rt, C, S, H, W):
    """Fused attention branch (mma.sync kernel) against torch fp64 on the same bf16-rounded operands: the only differences are
    fp32 accumulation order, the bf16 rounding of the intermediate and of the result."""
    B = 3
    bf = lambda t: t.to(torch.bfloat16).float()
    x = bf(_rand(B, C, S, H, W, seed=11, scale=2.0))
    wa = bf(_rand(C, C, 3, 1, 1, seed=12, scale=(2.0 / (3 * C)) ** 0.5 * 1.7))
    wb = bf(_rand(C, C, 1, 1, 1, seed=13, scale=(2.0 / C) ** 0.5 * 1.7))
    mid = F.relu(F.conv3d(x.double(), wa.double(), None, 1, (1, 0, 0)))
    mid_bf = mid.float().to(torch.bfloat16).double()
    ref = x.double() + F.relu(F.conv3d(mid_bf, wb.double()))
    out = rt.srd_attention(x.cuda(), wa.cuda(), wb.cuda()).cpu().double()
    # the intermediate may round to a neighbouring bf16 value (fp32 vs fp64 accumulation): one ulp of one term of the second sum
    tol = 2.0 ** -8 * (ref.abs().max().item() + mid.abs().max().item() * wb.abs().max().item())
    err = (out - ref).abs().max().item()
    assert err <= tol, (err, tol)
    # and the bulk must be exact to output rounding (half an ulp of bf16)
    frac_exact = ((out - ref).abs() <= 2.0 ** -8 * ref.abs().clamp_min(1e-3)).double().mean().item()
    assert frac_exact > 0.99, frac_exact


@pytest.mark.parametrize("C,k,is_max", [(8, 2, True), (16, 2, True), (32, 2, False), (32, 4, False), (32, 8, False)])
def test_pool_bf16_exact(rt, C, k, is_max):
    """(1,k,k) pooling on bf16 channels-last volumes (MaxPool3d of EFD, reference :387; AvgPool3d pyramid, :248-250): the maximum of
    bf16 values is exact; the average is one rounding of the fp32 mean."""
    import ctypes
    from dffinthewild_b200 import train as tr
    l = tr._lib() if hasattr(tr, "_lib") else rt.lib()
    BS, H, W = 6, 16, 48
    x = _rand(BS, H, W, C, seed=21, scale=3.0).to(torch.bfloat16).cuda()
    out = torch.empty((BS, H // k, W // k, C), dtype=torch.bfloat16, device="cuda")
    f = rt.lib().dff_pool3d
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 7 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    rt.check(f(x.data_ptr(), BS, H, W, C, k, 1 if is_max else 0, rt.BF16, out.data_ptr(), 0, torch.cuda.current_stream().cuda_stream))
    xr = x.float().view(BS, H // k, k, W // k, k, C)
    ref = xr.amax(dim=(2, 4)) if is_max else xr.mean(dim=(2, 4))
    if is_max:
        assert torch.equal(out.float(), ref)
    else:
        assert (out.float() - ref).abs().max().item() <= 2.0 ** -8 * ref.abs().max().item()


@pytest.mark.parametrize("r", [1, 2, 4, 8])
def test_depth_head(rt, r):
    from oracle import dff_oracle as O
    B, S, H, W = 2, 5, 32, 64
    cost = _rand(B, S, H // r, W // r, seed=20, scale=12.0)
    cost[0, 0, 0, 0] = 30.0  # softplus threshold branch
    for fd in (_rand(B, S, H, W, seed=21).abs() + 0.1, torch.linspace(0.1, 1.5, S).view(1, S, 1, 1)):
        ref = O.depth_head(cost.double(), fd.double(), (H, W))
        out = rt.depth_head(cost.cuda(), fd.cuda(), H, W).cpu().double()
        assert ((out - ref).abs() / ref.abs()).max().item() <= 2e-6


def test_depth_head_many_slices(rt):
    from oracle import dff_oracle as O
    cost, fd = _rand(1, 49, 8, 8, seed=22, scale=5.0), _rand(1, 49, 32, 32, seed=23).abs() + 0.05
    ref = O.depth_head(cost.double(), fd.double(), (32, 32))
    out = rt.depth_head(cost.cuda(), fd.cuda(), 32, 32).cpu().double()
    assert ((out - ref).abs() / ref.abs()).max().item() <= 2e-6


@pytest.mark.parametrize("tag", ["b1", "b2"])
def test_fov_warp_matches_reference_golden(rt, tag):
    g = golden("g5_fov_warp_%s.npz" % tag)
    x, alpha, fov = (torch.from_numpy(g[k]).cuda() for k in ("x", "alpha", "fov"))
    out, flow = rt.fov_warp(x, alpha, fov)
    assert np.abs(flow.cpu().numpy() - g["flow"]).max() <= 2e-5
    assert np.abs(out.cpu().numpy() - g["out"]).max() <= 2e-4   # bilinear weights amplify 1-ulp coordinate noise
    out0, _ = rt.fov_warp(x, None, fov)
    from oracle import dff_oracle as O
    ref0, _ = O.fov_warp(torch.from_numpy(g["x"]), torch.zeros(1, 3, 1, 1), torch.from_numpy(g["fov"]))
    assert (out0.cpu() - ref0).abs().max().item() <= 2e-4


@pytest.mark.parametrize("H,W,scale", [(96, 128, 0.05), (70, 132, 0.3), (64, 1024, 0.02), (48, 100, 0.05), (40, 36, 1.5), (32, 50, 0.05)])
def test_fov_warp_staged_rows_kernel(rt, H, W, scale):
    """The shared-memory-staged FOV warp (W % 4 == 0: windows of source rows per block of 16 output rows; scale 0.3 / 1.5 push
    windows past the stage so that items fall back to direct sampling) and the gather kernel (W % 4 != 0) against the oracle's
    grid_sample restatement (End_to_End/End_to_End.py:106-134), shifts of several pixels, flow output included."""
    from oracle import dff_oracle as O
    B, C, S = 2, 3, 3
    g = torch.Generator().manual_seed(H * 1000 + W)
    x = torch.rand(B, C, S, H, W, generator=g) * 2 - 1
    alpha = (torch.rand(B, 3, S, 1, 1, generator=g) - 0.5) * torch.tensor([2 * scale, 12.0, 9.0]).view(1, 3, 1, 1, 1)
    fov = 1.0 + 0.02 * torch.arange(S, dtype=torch.float32).view(1, 1, S, 1, 1).repeat(B, 1, 1, 1, 1)
    ref, rflow = O.fov_warp(x, alpha, fov)
    out, flow = rt.fov_warp(x.cuda(), alpha.cuda(), fov.cuda())
    assert (flow.cpu() - rflow).abs().max().item() <= 1e-4 * max(1.0, rflow.abs().max().item())
    # bilinear weights amplify 1-ulp coordinate noise (coordinates up to W): compare where the sample position is well conditioned
    d = (out.cpu() - ref).abs()
    assert d.max().item() <= 2e-3 and d.mean().item() <= 2e-5, (d.max().item(), d.mean().item())


def test_layout_round_trip(rt):
    x = _rand(2, 3, 4, 8, 12, seed=30).cuda()
    cl = rt.to_channels_last(x, 4)
    assert cl.shape == (2, 4, 8, 12, 4) and float(cl[..., 3].abs().max()) == 0.0
    assert torch.equal(rt.from_channels_last(cl, 3), x)
    assert torch.equal(cl[..., :3].permute(0, 4, 1, 2, 3), x)


@pytest.mark.parametrize("fd_kind", ["scalars", "tiled", "strided"])
@pytest.mark.parametrize("B,S,H,W", [(2, 5, 32, 32), (1, 10, 64, 96), (1, 3, 16, 16)])
def test_depth_heads4_quad_kernel(rt, fd_kind, B, S, H, W):
    """All four heads in one launch (C-ABI dff_depth_heads4): the bf16 mode's four-pixels-per-thread kernel (fixed 1/8, 1/4, 1/2, 1/1
    pyramid, replicate-edge column loads) and the fp32 mode's kernel against the oracle's depth head (reference :92-98, 118-136)."""
    import ctypes
    from oracle import dff_oracle as O
    costs = [_rand(B, S, H // r, W // r, seed=70 + r, scale=14.0) for r in (8, 4, 2, 1)]
    costs[3][0, 0, 0, 0] = 31.0    # softplus threshold branch
    if fd_kind == "scalars":
        fd = torch.linspace(0.1, 1.5, S).view(1, S, 1, 1).expand(B, S, 1, 1).contiguous().cuda().expand(B, S, H, W)
    elif fd_kind == "tiled":
        fd = (_rand(B, S, H, W, seed=75).abs() + 0.1).cuda()
    else:
        fd = (_rand(B, S, H, W + 3, seed=76).abs() + 0.1).cuda()[..., 1:W + 1]     # unaligned rows: the generic-stride variant
    f = rt.lib().dff_depth_heads4
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)] + [ctypes.c_int] * 4 + \
                 [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    dc = [c.cuda() for c in costs]
    cp = (ctypes.c_void_p * 4)(*[c.data_ptr() for c in dc])
    strides = (ctypes.c_int64 * 4)(*fd.stride())
    refs = [O.depth_head(c.double(), fd.cpu().double(), (H, W)) for c in costs]
    for fast, tol in ((1, 2e-5), (0, 2e-6)):
        outs = [torch.full((B, H, W), float("nan"), device="cuda") for _ in range(4)]
        op = (ctypes.c_void_p * 4)(*[o.data_ptr() for o in outs])
        rt.check(f(cp, fd.data_ptr(), strides, B, S, H, W, op, fast, 0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        for o, r, k in zip(outs, refs, range(4)):
            err = ((o.cpu().double() - r).abs() / r.abs()).max().item()
            assert err <= tol, (fast, k, err)


@pytest.mark.parametrize("C", [8, 16, 32])
def test_fov_warp_channels_last_bf16(rt, C):
    """bf16 channels-last FOV warp (16-byte pieces) against the planar fp32 kernel — itself pinned to the reference goldens above —
    on the same bf16-rounded volume: identical geometry, one bf16 rounding of the blended value."""
    import ctypes
    B, S, H, W = 2, 10, 24, 40
    g = torch.Generator().manual_seed(91)
    x = (torch.rand(B, C, S, H, W, generator=g) * 2 - 1).bfloat16().float().cuda()
    alpha = (torch.randn(B, 3, S, generator=g) * torch.tensor([0.01, 1.5, 1.5]).view(1, 3, 1)).cuda().contiguous()
    fov = (torch.linspace(1.02, 1.0, S).view(1, S).expand(B, S) + 0.003 * torch.arange(B).view(B, 1)).cuda().contiguous()
    ref, _ = rt.fov_warp(x, alpha.view(B, 3, S, 1, 1), fov.view(B, 1, S, 1, 1), want_flow=False)
    xcl = rt.to_channels_last(x, C, True)
    out = torch.empty_like(xcl)
    rt.check(rt.lib().dff_fov_warp_cl(xcl.data_ptr(), alpha.data_ptr(), fov.data_ptr(), B, C, S, H, W, out.data_ptr(), rt.BF16, 0,
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    got = rt.from_channels_last(out, C)
    assert (got - ref).abs().max().item() <= 2.0 ** -8 * ref.abs().max().item() + 1e-6


@pytest.mark.parametrize("C,BS,H,W", [(32, 6, 16, 48), (8, 3, 8, 8), (16, 2, 24, 40)])
def test_avgpool_pyramid_one_pass(rt, C, BS, H, W):
    """The three average pools of hourglassup in one pass (reference :248-250): each level = one bf16 rounding of the fp32 mean."""
    import ctypes
    x = _rand(BS, H, W, C, seed=23, scale=3.0).to(torch.bfloat16).cuda()
    outs = [torch.full((BS, H // k, W // k, C), float("nan"), dtype=torch.bfloat16, device="cuda") for k in (2, 4, 8)]
    f = rt.lib().dff_avgpool_pyramid
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_void_p]
    rt.check(f(x.data_ptr(), BS, H, W, C, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), 0,
               torch.cuda.current_stream().cuda_stream))
    for k, o in zip((2, 4, 8), outs):
        ref = x.float().view(BS, H // k, k, W // k, k, C).mean(dim=(2, 4))
        assert (o.float() - ref).abs().max().item() <= 2.0 ** -8 * ref.abs().max().item(), k
