"""The reference's depth metrics (`metrics.py:90-133`, identical in train_codes/ and Depth_Estimation_Test/) and the output image of
its eval loop (`Depth_Estimation_Test/test.py:123-140`) computed on the GPU from the forward's own output tensors.

Same function names and argument order as the reference module, but the arguments are CUDA tensors and the results stay on the
device (0-d tensors; `.item()` when the caller wants the number), so an eval loop does not synchronise per stack:

    from dffinthewild_b200 import metrics as M
    fig = M.depth_metrics(pred3, gt, mask)            # dict of (B,) tensors: abs_rel, sq_rel, mse, mae, rmse, rmse_log, accuracy_1..3
    M.mask_abs_rel(pred3[0], gt[0], mask[0])          # the reference's signature, one map

`est` may be the padded (H, W) map the network returns: metrics are taken over the `[:gt.shape[-2], :gt.shape[-1]]` crop exactly as
test.py:125 crops before calling metrics.py.
"""
import ctypes

import torch

from . import runtime as rt

NAMES = ("abs_rel", "sq_rel", "mse", "mae", "rmse", "rmse_log", "accuracy_1", "accuracy_2", "accuracy_3", "mse_w_conf", "mae_w_conf",
         "count")
_P = ctypes.c_void_p


def _declare(l):
    if getattr(l, "_dff_metrics_declared", False):
        return
    c = ctypes
    l.dff_depth_metrics_scratch_bytes.restype = c.c_size_t
    l.dff_depth_metrics_scratch_bytes.argtypes = [c.c_int]
    l.dff_depth_metrics.restype = c.c_int
    l.dff_depth_metrics.argtypes = [_P, _P, _P, _P] + [c.c_int] * 5 + [_P, _P, c.c_int, _P]
    l.dff_depth_to_jet.restype = c.c_int
    l.dff_depth_to_jet.argtypes = [_P] + [c.c_int] * 5 + [c.c_float, c.c_float, _P, _P, c.c_int, _P]
    l._dff_metrics_declared = True


def _p(t):
    return _P(t.data_ptr()) if t is not None else _P(0)


def _maps(t):
    return t if t.dim() == 3 else t.unsqueeze(0)


def depth_metrics(est, gt, mask=None, conf=None):
    """All figures of metrics.py:90-133 for B maps in one pass: est (B,H,W) | (H,W); gt / mask / conf (B,Hc,Wc) with Hc <= H, Wc <= W."""
    rt._require_cuda(est, "est")
    rt._require_cuda(gt, "gt")
    l = rt.lib()
    _declare(l)
    est, gt = _maps(est).float().contiguous(), _maps(gt).float().contiguous()
    B, H, W = est.shape
    _, Hc, Wc = gt.shape
    m8 = None
    if mask is not None:
        mask = _maps(mask)
        m8 = (mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else (mask != 0).to(torch.uint8)).contiguous()
    cf = _maps(conf).float().contiguous() if conf is not None else None
    dev = est.device
    out = torch.empty((B, 12), dtype=torch.float32, device=dev)
    scratch = torch.empty(l.dff_depth_metrics_scratch_bytes(B), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rt.check(l.dff_depth_metrics(_p(est), _p(gt), _p(m8), _p(cf), B, H, W, Hc, Wc, _p(out), _p(scratch), dev.index,
                                     _P(torch.cuda.current_stream(dev).cuda_stream)))
    return {n: out[:, i] for i, n in enumerate(NAMES)}


def _one(name, est, gt, mask, conf=None):
    return depth_metrics(est, gt, mask, conf)[name][0]


def mask_abs_rel(est_depth, gt_depth, mask):
    return _one("abs_rel", est_depth, gt_depth, mask)


def mask_sq_rel(est_depth, gt_depth, mask):
    return _one("sq_rel", est_depth, gt_depth, mask)


def mask_mse(est_depth, gt_depth, mask):
    return _one("mse", est_depth, gt_depth, mask)


def mask_mae(est_depth, gt_depth, mask):
    return _one("mae", est_depth, gt_depth, mask)


def mask_rmse(est_depth, gt_depth, mask):
    return _one("rmse", est_depth, gt_depth, mask)


def mask_rmse_log(est_depth, gt_depth, mask):
    return _one("rmse_log", est_depth, gt_depth, mask)


def mask_accuracy_k(est_depth, gt_depth, k, mask):
    if k not in (1, 2, 3):
        raise rt.DffError("dff_b200: mask_accuracy_k is computed for k = 1, 2, 3 (the values test.py uses)")
    return _one("accuracy_%d" % k, est_depth, gt_depth, mask)


def mask_mse_w_conf(est_depth, gt_depth, conf, mask):
    return _one("mse_w_conf", est_depth, gt_depth, mask, conf)


def mask_mae_w_conf(est_depth, gt_depth, conf, mask):
    return _one("mae_w_conf", est_depth, gt_depth, mask, conf)


def depth_to_jet(est, crop_hw, min_depth, max_depth):
    """test.py:123-140: crop, normalise to [min_depth, max_depth], matplotlib 'jet' -> (B,Hc,Wc,3) uint8 on the device."""
    rt._require_cuda(est, "est")
    l = rt.lib()
    _declare(l)
    est = _maps(est).float().contiguous()
    B, H, W = est.shape
    Hc, Wc = crop_hw
    dev = est.device
    out = torch.empty((B, Hc, Wc, 3), dtype=torch.uint8, device=dev)
    lut = torch.empty(768, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rt.check(l.dff_depth_to_jet(_p(est), B, H, W, Hc, Wc, float(min_depth), float(max_depth), _p(out), _p(lut), dev.index,
                                    _P(torch.cuda.current_stream(dev).cuda_stream)))
    return out
