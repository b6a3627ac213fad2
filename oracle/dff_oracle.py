"""CPU oracle — TEST INFRASTRUCTURE ONLY (never imported by the product path).

A functional restatement of the reference's depth-from-focus forward pass, written from the dataflow of
`/root/reference/train_codes/Depth_Estimation_Network.py` (T-DEN) and its inference twin
`Depth_Estimation_Test/Depth_Estimation_Network.py` (E-DEN).  It takes a plain `state_dict` (the reference's
384-key layout) plus inputs and evaluates the network with `torch.nn.functional` primitives on the CPU in
fp32 or fp64.  The arithmetic primitives (conv3d, conv_transpose3d, batch_norm, pooling, bilinear
interpolate, softplus, grid_sample) live in PyTorch, a third-party dependency the reference pins as
torch==1.6.0 (README.md:16) and that is installed here as 2.11.0; the kernel-level tests check each primitive
against the same torch CPU call in fp64 (`tests/test_gpu_ops.py`).

Pinning: the reference ships no tests, golden vectors or checkpoints, so this oracle is
pinned against the *live reference module* imported from `/root/reference` in the build container:
`oracle/gen_golden.py` runs the unmodified reference and this restatement on the same seeded inputs, asserts
bit-equality, and commits the reference outputs under `tests/golden/`.  `tests/test_oracle_golden.py`
re-checks the restatement against those committed vectors wherever the tests run.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py` (cpu_baseline / `--impl reference`) may import this.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
UP = dict(stride=(1, 2, 2), padding=1, output_padding=(0, 1, 1))


class Ctx:
    """Evaluation context: weights, BN mode, optional recording of every intermediate by name."""

    def __init__(self, sd, train=False, prefix="DFF_net.", record=None, update_stats=False):
        self.sd, self.train, self.p, self.rec, self.update = sd, train, prefix, record, update_stats

    def w(self, name):
        return self.sd[self.p + name]

    def keep(self, name, t):
        if self.rec is not None:
            self.rec[name] = t
        return t


def _bn(c, x, name):
    """nn.BatchNorm3d (T-DEN:355): eval = running stats; train = biased batch variance, momentum 0.1."""
    g, b = c.w(name + ".weight"), c.w(name + ".bias")
    rm, rv = c.w(name + ".running_mean"), c.w(name + ".running_var")
    if not c.train:
        return F.batch_norm(x, rm, rv, g, b, False, BN_MOMENTUM, BN_EPS)
    if c.update:
        out = F.batch_norm(x, rm, rv, g, b, True, BN_MOMENTUM, BN_EPS)
        c.sd[c.p + name + ".num_batches_tracked"] += 1
        return out
    return F.batch_norm(x, None, None, g, b, True, BN_MOMENTUM, BN_EPS)


def conv(c, x, name, stride=1, pad=1, dil=1):
    return F.conv3d(x, c.w(name + ".weight"), None, stride, pad, dil)


def conv_bn(c, x, name, stride=1, pad=1, dil=1):
    """convbn_3d (T-DEN:352-355): `name.0` conv, `name.1` BN."""
    return _bn(c, conv(c, x, name + ".0", stride, pad, dil), name + ".1")


def up_bn(c, x, name):
    """ConvTranspose3d k3 s(1,2,2) p1 op(0,1,1) + BN (T-DEN:43-50)."""
    return _bn(c, F.conv_transpose3d(x, c.w(name + ".0.weight"), None, **UP), name + ".1")


def srd(c, x, name):
    """Feature_Extraction / SRD (T-DEN:394-407, resnet_block_2d :361-370)."""
    fm = name + ".Focus_Measure.conv"
    t = F.relu(conv_bn(c, x, fm + ".0", 1, (0, 1, 1)))
    f = F.relu(x + conv_bn(c, t, fm + ".2", 1, (0, 1, 1)))
    a = F.relu(conv(c, f, name + ".N_ch_attention.0", 1, (1, 0, 0)))
    a = F.relu(conv(c, a, name + ".N_ch_attention.2", 1, 0))
    return f + a


def efd(c, x, name):
    """res_stride_conv_3d / EFD (T-DEN:383-392)."""
    a = conv_bn(c, x, name + ".stride_conv", (1, 2, 2), 1)
    b = conv_bn(c, F.max_pool3d(x, (1, 2, 2), (1, 2, 2)), name + ".max_pooling.1", 1, 1)
    return F.relu(a + b)


def pyramid(c, x, name):
    """hourglassup (T-DEN:247-273)."""
    def tower(t, n0, n1):
        r = F.relu(conv_bn(c, F.relu(conv_bn(c, t, n0 + ".0")), n0 + ".2"))
        return conv_bn(c, F.relu(conv_bn(c, r, n1 + ".0")), n1 + ".2") + r

    x32 = F.avg_pool3d(x, (1, 8, 8), (1, 8, 8))
    x16 = F.avg_pool3d(x, (1, 4, 4), (1, 4, 4))
    x8 = F.avg_pool3d(x, (1, 2, 2), (1, 2, 2))
    x8 = c.keep(name + ".x8", tower(x8, name + ".dres8_0", name + ".dres8_1"))
    x16 = c.keep(name + ".x16", tower(x16, name + ".dres16_0", name + ".dres16_1"))
    x32 = c.keep(name + ".x32", tower(x32, name + ".dres32_0", name + ".dres32_1"))
    c1 = conv(c, x8, name + ".conv1", (1, 2, 2), 1)
    c1 = F.relu(conv_bn(c, torch.cat((c1, x16), 1), name + ".combine1.0"))
    c2 = F.relu(conv_bn(c, c1, name + ".conv2.0"))
    c3 = conv(c, c2, name + ".conv3", (1, 2, 2), 1)
    c3 = F.relu(conv_bn(c, torch.cat((c3, x32), 1), name + ".combine2.0"))
    c4 = F.relu(conv_bn(c, c3, name + ".conv4.0"))
    c8 = F.relu(up_bn(c, c4, name + ".conv8") + conv_bn(c, c2, name + ".redir2", 1, 0))
    c9 = F.relu(up_bn(c, c8, name + ".conv9") + conv_bn(c, x8, name + ".redir1", 1, 0))
    return c9


def hourglass(c, x, name, presqu, postsqu):
    """hourglass (T-DEN:302-321) -> (out, pre_1)."""
    pre_1 = F.relu(conv_bn(c, x, name + ".conv0.0"))
    out = F.relu(conv_bn(c, pre_1, name + ".conv1.0", (1, 2, 2), 1))
    pre = conv_bn(c, out, name + ".conv2")
    pre = F.relu(pre + postsqu) if postsqu is not None else F.relu(pre)
    out = F.relu(conv_bn(c, pre, name + ".conv3.0", (1, 2, 2), 1))
    out = F.relu(conv_bn(c, out, name + ".conv4.0"))
    out = F.relu(up_bn(c, out, name + ".conv5") + (presqu if presqu is not None else pre))
    out = up_bn(c, out, name + ".conv6")
    return out, pre_1


def depth_head(cost, focus_dists, size):
    """Bilinear upsample (align_corners=False) -> softplus + 1e-6 -> normalise over S -> E[fd] (T-DEN:92-98)."""
    if tuple(cost.shape[-2:]) != tuple(size):
        cost = F.interpolate(cost, size=list(size), mode="bilinear", align_corners=False)
    p = F.softplus(cost) + 1e-6
    p = p / p.sum(dim=1, keepdim=True)
    return torch.sum(focus_dists * p, dim=1)


def dff_forward(sd, FS, focus_dists, train=False, prefix="DFF_net.", record=None, update_stats=False):
    """DFF_net.forward (T-DEN:77-137).  FS (B,3,S,H,W); focus_dists broadcastable to (B,S,H,W)."""
    c = Ctx(sd, train, prefix, record, update_stats)
    H, W = FS.shape[-2:]
    x = F.relu(conv_bn(c, FS, "FM_measure.Focus_extraction.0", 1, (0, 8, 8), (1, 2, 2)))
    v1 = c.keep("V1", srd(c, x, "FM_measure.Focus_extraction.2"))
    v2 = c.keep("V2", srd(c, efd(c, v1, "FM_conv1.0"), "FM_conv1.1"))
    v3 = c.keep("V3", srd(c, efd(c, v2, "FM_conv2.0"), "FM_conv2.1"))
    vol = c.keep("FS_volume", pyramid(c, v3, "SPP_module"))

    cm = conv(c, F.relu(conv_bn(c, vol, "confidence.0")), "confidence.2")
    cm = c.keep("cost_mid", cm.squeeze(1))
    mid_out = depth_head(cm, focus_dists, (H, W))

    x = F.relu(conv_bn(c, vol, "dres0.0"))
    x = F.relu(conv_bn(c, x, "dres0.2"))
    x = c.keep("x", up_bn(c, x, "deconv_1"))
    out, pre = hourglass(c, torch.cat([x, v3], 1), "dres2", None, None)
    out_in = x + out
    cost1 = c.keep("cost1", conv(c, out_in, "classif1.0", 1, 0).squeeze(1))

    out2 = up_bn(c, out_in, "deconv_2")
    out, pre = hourglass(c, torch.cat([out2, v2], 1), "dres3", pre, out)
    out_in = out2 + out
    cost2 = c.keep("cost2", conv(c, out_in, "classif2.0", 1, 0).squeeze(1))

    out2 = up_bn(c, out_in, "deconv_3")
    out, _ = hourglass(c, torch.cat([out2, v1], 1), "dres4", pre, out)
    out = out2 + out
    cost3 = c.keep("cost3", conv(c, out, "classif3.0", 1, 0).squeeze(1))

    pred1 = depth_head(cost1, focus_dists, (H, W))
    pred2 = depth_head(cost2, focus_dists, (H, W))
    pred3 = depth_head(cost3, focus_dists, (H, W))
    return mid_out, pred1, pred2, pred3


# ----------------------------------------------------------------------------------------------------------
# End-to-End alignment network (End_to_End/End_to_End.py, "E2E")
# ----------------------------------------------------------------------------------------------------------

def fov_warp(x, alpha, fovs):
    """FlowNetwork.FOV_warp (E2E:106-134) in closed form for B == 1 semantics per sample.

    x (B,C,S,H,W); alpha (B,3,S,1,1) or (1,3,1,1) zeros; fovs (B,1,S,1,1) -> (warped x, flow (B,2,S,H,W)).
    f = fovs + alpha[:,0]; flow_x = (W//2)(f-1)*lin(-1,1,W) + alpha[:,1]; flow_y likewise with H and alpha[:,2];
    sample slice s bilinearly at (x-flow_x, y-flow_y), zeros padding, align_corners=True.  The reference's
    batch>1 broadcast quirk is reproduced by following its tensor algebra literally.
    """
    B, C, S, H, W = x.shape
    f = alpha[:, 0, :, :] + fovs                      # (B,S,1,1)+(B,1,S,1,1) -> (B,B,S,1,1) exactly as E2E:112
    lx = torch.linspace(-1, 1, steps=W, dtype=x.dtype).view(1, 1, 1, W).expand(B, S, H, W)
    ly = torch.linspace(-1, 1, steps=H, dtype=x.dtype).view(1, 1, H, 1).expand(B, S, H, W)
    fx = (W // 2) * (f[:, 0] - 1) * lx + alpha[:, 1, :, :]
    fy = (H // 2) * (f[:, 0] - 1) * ly + alpha[:, 2, :, :]
    flow = torch.stack((fx, fy), 1).to(x.dtype)
    gx = torch.arange(W, dtype=x.dtype).view(1, 1, 1, W) - fx
    gy = torch.arange(H, dtype=x.dtype).view(1, 1, H, 1) - fy
    gz = torch.arange(S, dtype=x.dtype).view(1, S, 1, 1).expand(B, S, H, W)
    grid = torch.stack((2.0 * gx / max(W - 1, 1) - 1.0, 2.0 * gy / max(H - 1, 1) - 1.0,
                        2.0 * gz / max(S - 1, 1) - 1.0), -1)
    return F.grid_sample(x, grid, align_corners=True), flow


def _res2d_of(c, x, name, stride):
    """resnet_block_2d_OF (E2E:135-145)."""
    t = F.relu(conv_bn(c, x, name + ".conv.0", (1, stride, stride), (0, 1, 1)))
    t = conv_bn(c, t, name + ".conv.2", 1, (0, 1, 1))
    return F.relu(conv(c, x, name + ".feature", (1, stride, stride), 0) + t)


def _align_head(c, vol, name):
    """conv1/conv2/conv3 of FlowNetwork (E2E:33-61): 3x(conv1x3x3+BN+ReLU), biased conv -> 3ch, per-slice mean."""
    t = vol
    for i in (0, 2, 4):
        t = F.relu(conv_bn(c, t, "%s.%d" % (name, i), 1, (0, 1, 1)))
    t = F.conv3d(t, c.w(name + ".6.weight"), c.w(name + ".6.bias"), 1, (0, 1, 1))
    if t.shape[2] != 10:
        raise RuntimeError("FlowNetwork hard-wires S == 10 (AdaptiveAvgPool3d((10,1,1)), E2E:40)")
    return F.adaptive_avg_pool3d(t, (10, 1, 1))      # per-slice spatial mean


def _pair_volume(fe, flow):
    """The FE*_copy builders (E2E:71-76): [last-slice features | slice-i features | flow]."""
    last = fe[:, :, -1:, :, :].expand_as(fe)
    return torch.cat((last, fe, flow), 1).contiguous()


def flow_forward(sd, FS, fovs, prefix="optical_flow_aggregation.", train=False, return_alpha=False):
    """FlowNetwork.forward (E2E:63-104): returns the aligned focal stack [and the accumulated alpha (B,3,S,1,1)]."""
    c = Ctx(sd, train, prefix)
    fe1 = _res2d_of(c, _res2d_of(c, FS, "OF_feature.0", 1), "OF_feature.1", 1)
    fe2 = _res2d_of(c, _res2d_of(c, fe1, "OF_feature1.0", 2), "OF_feature1.1", 1)
    fe3 = _res2d_of(c, _res2d_of(c, fe2, "OF_feature2.0", 2), "OF_feature2.1", 1)
    zero = torch.zeros((1, 3, 1, 1), dtype=FS.dtype)
    fe3, flow = fov_warp(fe3, zero, fovs)
    alpha = _align_head(c, _pair_volume(fe3, flow), "conv1")
    alpha = torch.cat((0.001 * alpha[:, :1], alpha[:, 1:]), 1)
    fe2, flow = fov_warp(fe2, alpha, fovs)
    na = _align_head(c, _pair_volume(fe2, flow), "conv2")
    alpha = torch.cat((0.001 * na[:, :1], na[:, 1:]), 1) + alpha
    fe1, flow = fov_warp(fe1, alpha, fovs)
    na = _align_head(c, _pair_volume(fe1, flow), "conv3")
    alpha = torch.cat((0.001 * na[:, :1], na[:, 1:]), 1) + alpha
    out, _ = fov_warp(FS, alpha, fovs)
    return (out, alpha) if return_alpha else out


def e2e_forward(sd, FS, focus_dists, fovs):
    """End_to_End.Network.forward (E2E:14-17, 259): 5-tuple, 5th = aligned stack."""
    warped = flow_forward(sd, FS, fovs)
    return dff_forward(sd, warped, focus_dists) + (warped,)


# ----------------------------------------------------------------------------------------------------------
# training recipe (train_codes/train_code_Defocus.py:17-19, 34-38, 160-165)
# ----------------------------------------------------------------------------------------------------------

LOSS_WEIGHTS = (0.3, 0.5, 0.7, 1.0)   # (mid, pred1, pred2, pred3)


def defocus_loss(outs, gt, mask):
    """0.5*L1 + 0.7*L2 + 1.0*L3 + 0.3*Lmid, L = mean squared error over masked pixels."""
    mid, p1, p2, p3 = (F.mse_loss(o[mask], gt[mask]) for o in outs)
    return 0.5 * p1 + 0.7 * p2 + 1.0 * p3 + 0.3 * mid      # the reference's summation order


# ----------------------------------------------------------------------------------------------------------
# metrics (train_codes/metrics.py:90-97, 41-61)
# ----------------------------------------------------------------------------------------------------------

def mask_mse(est, gt, mask):
    import numpy as np
    return float(np.mean((est[mask] - gt[mask]) ** 2))


def mask_abs_rel(est, gt, mask):
    import numpy as np
    return float(np.mean(np.abs(est[mask] - gt[mask]) / gt[mask]))


def depth_metrics(est, gt, mask, conf=None):
    """Every masked figure of metrics.py:90-133 (abs_rel :90-91, sq_rel :93-94, mse :96-97, mae :99-100, rmse :102-103,
    rmse_log :105-109, accuracy_k :112-121, *_w_conf :123-127) in the reference's own numpy arithmetic."""
    import numpy as np
    e, g = est[mask], gt[mask]
    out = {"abs_rel": np.mean(np.abs(g - e) / g), "sq_rel": np.mean(np.power(g - e, 2) / g), "mse": np.mean(np.power(g - e, 2)),
           "mae": np.mean(np.abs(g - e)), "rmse": np.sqrt(np.mean(np.power(e - g, 2))),
           "rmse_log": np.sqrt(np.mean(np.power(np.log(g) - np.log(e), 2)))}
    th = np.maximum(e / g, g / e)
    for k in (1, 2, 3):
        out["accuracy_%d" % k] = np.sum(np.where(th < (1.25 ** k), 1, 0)) / np.sum(mask)
    if conf is not None:
        c = conf[mask]
        out["mse_w_conf"] = np.sum(c * np.power(g - e, 2)) / np.sum(c)
        out["mae_w_conf"] = np.sum(c * np.abs(g - e)) / np.sum(c)
    return {k: float(v) for k, v in out.items()}


def bumpiness(gt, est, mask, clip=0.05, factor=100):
    """get_bumpiness (metrics.py:41-61) with scikit-image's Scharr filters restated via scipy.ndimage
    (skimage 0.19.2 is not installed and not vendored: bumpiness parity is unpinned)."""
    import numpy as np
    from scipy import ndimage as ndi
    k = np.array([[3, 10, 3], [0, 0, 0], [-3, -10, -3]], dtype=np.float64) / 16.0

    def sv(a):  # vertical-edge Scharr (derivative along columns)
        return ndi.convolve(a.astype(np.float64), k.T, mode="reflect")

    def sh(a):
        return ndi.convolve(a.astype(np.float64), k, mode="reflect")

    d = np.asarray(gt, dtype=np.float64) - np.asarray(est, dtype=np.float64)
    dx, dy = sv(d), sh(d)
    dxx, dxy, dyy, dyx = sv(dx), sh(dx), sh(dy), sv(dy)
    b = np.sqrt(dxx ** 2 + dxy ** 2 + dyy ** 2 + dyx ** 2)
    b = np.clip(b, 0, clip)
    return float(np.mean(b[mask]) * factor)
