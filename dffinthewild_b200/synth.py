"""Deterministic synthetic weights and inputs for tests/, bench.py and oracle/gen_golden.py (generators only: no arithmetic of the path).

Everything is drawn from numpy's PCG64 (stable across numpy versions), keyed by name, so the same tensors can be
rebuilt on the GPU box without shipping a 16 MB state_dict.  Input recipes follow SURVEY.md §8(d):
FS ~ U(-1,1) with the padded rows/cols set to -1 (reference `test_Dataloader.py:128-140`), DDFF focus distances
linspace(0.28248, 0.020177, 10) (`test_Dataloader.py:105-113`), DefocusNet [0.1,0.15,0.3,0.7,1.5]
(`train_Dataloader.py:89`).
"""
import zlib

import numpy as np
import torch

DDFF_FD = (0.28248, 0.020177)
DEFOCUS_FD = (0.1, 0.15, 0.3, 0.7, 1.5)
# conv std = CONV_GAIN/sqrt(fan_in): 1.2 keeps activations O(1-30) and the pre-softplus costs at |c| ~ 10-40 through all 70
# layers (sqrt(2) blows them up to ~1e3, 1.0 collapses the depth maps to the mean focus distance) — "trained-like".
CONV_GAIN = 1.2


def _rng(seed, name):
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))


def synthetic_state(template, seed=1):
    """A 'trained-like' state_dict with the template's keys/shapes: fan-in-scaled conv weights and non-trivial
    BN affine + running statistics, so activations stay O(1) and every BN term is exercised."""
    out = {}
    for k, v in template.items():
        r = _rng(seed, k)
        shp = tuple(v.shape)
        if k.endswith("num_batches_tracked"):
            t = torch.tensor(7, dtype=torch.int64)
        elif k.endswith("running_mean"):
            t = torch.from_numpy(r.normal(0, 0.2, shp).astype(np.float32))
        elif k.endswith("running_var"):
            t = torch.from_numpy(r.uniform(0.5, 1.5, shp).astype(np.float32))
        elif v.dim() == 1 and k.endswith("weight"):      # BN gamma
            t = torch.from_numpy(r.uniform(0.6, 1.4, shp).astype(np.float32))
        elif v.dim() == 1:                                 # BN beta / conv bias
            t = torch.from_numpy(r.normal(0, 0.1, shp).astype(np.float32))
        else:                                              # conv (Cout,Cin,kd,kh,kw) / deconv (Cin,Cout,kd,kh,kw)
            fan_in = int(np.prod(shp[1:]))
            t = torch.from_numpy(r.normal(0, CONV_GAIN / np.sqrt(max(fan_in, 1)), shp).astype(np.float32))
        out[k] = t
    return out


def focal_stack(B, S, H, W, seed=0, valid_hw=None):
    """FS (B,3,S,H,W) fp32 in [-1,1]; rows >= valid_hw[0] / cols >= valid_hw[1] are -1 like the dataloader pad."""
    r = _rng(seed, "FS")
    fs = r.uniform(-1, 1, (B, 3, S, H, W)).astype(np.float32)
    if valid_hw is not None:
        fs[..., valid_hw[0]:, :] = -1.0
        fs[..., :, valid_hw[1]:] = -1.0
    return torch.from_numpy(fs)


def focus_dists(B, S, H, W, kind="ddff", tiled=True):
    if kind == "ddff":
        fd = np.linspace(DDFF_FD[0], DDFF_FD[1], S, dtype=np.float64).astype(np.float32)
    elif kind == "defocus":
        base = np.asarray(DEFOCUS_FD, dtype=np.float32)
        fd = base[:S] if S <= len(base) else np.linspace(0.1, 1.5, S, dtype=np.float32)
    else:
        raise ValueError(kind)
    t = torch.from_numpy(fd).view(1, S, 1, 1)
    return t.expand(B, S, H, W).contiguous() if tiled else t.expand(B, S, 1, 1).contiguous()


def fovs(B, S):
    """Relative FOV per slice, (B,1,S,1,1), like `Test_dataloader.Real_Scenes` (:44-48): 1.02 ... 1.0."""
    return torch.linspace(1.02, 1.0, S).view(1, 1, S, 1, 1).expand(B, 1, S, 1, 1).contiguous()


def gt_and_mask(B, H, W, seed=0, lo=0.1, hi=1.5):
    r = _rng(seed, "gt")
    gt = r.uniform(lo, hi, (B, H, W)).astype(np.float32)
    gt[r.uniform(0, 1, (B, H, W)) < 0.1] = 0.0
    gt = torch.from_numpy(gt)
    return gt, gt > 0
