"""Achieved HBM GB/s of the bandwidth kernels on their algorithmic bytes (north_star (c)): FOV warp (End_to_End.py:106-134) at the
C4 shape, the depth heads (Depth_Estimation_Network.py:92-98,118-136) at the C2 shape, uint8 input staging.
    python tools/bw_kernels.py [--out profiles/r2_bw_kernels.json]"""
import argparse, ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dffinthewild_b200 import runtime as rt

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
a = ap.parse_args()
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6451.2
l = rt.lib()
dev = torch.device("cuda", 0)
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {"hbm_peak_gbs": peak, "kernels": {}}

def rec(name, ms, nbytes, note):
    gbs = nbytes / (ms / 1e3) / 1e9
    res["kernels"][name] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": gbs, "frac_of_hbm_peak": gbs / peak, "note": note}

# FOV warp of the focal stack itself (reference layout), 8 stacks of the C4 shape so the tensors exceed L2 (8 x 47 MB in, same out)
B, C, S, H, W = 8, 3, 10, 512, 768
x = torch.rand(B, C, S, H, W, device=dev) * 2 - 1
alpha = (torch.randn(B, 3, S, device=dev) * torch.tensor([0.002, 1.5, 1.5], device=dev).view(1, 3, 1)).contiguous()
fov = torch.linspace(1.02, 1.0, S, device=dev).view(1, S).expand(B, S).contiguous()
out = torch.empty_like(x)
ms = timed(lambda: rt.check(l.dff_fov_warp(x.data_ptr(), alpha.data_ptr(), fov.data_ptr(), B, C, S, H, W, out.data_ptr(), None, 0, st())))
rec("fov_warp (B,3,S,H,W) fp32, 8 x C4", ms, 2 * 4 * x.numel(), "1 read + 1 write of the stack; analytic coordinates, no grid tensor")
# channels-last feature volumes as FlowNetwork warps them (32 ch @1/4, 16 @1/2, 8 @1/1), bf16
for Cc, r in ((32, 4), (16, 2), (8, 1)):
    xc = torch.rand(B, S, H // r, W // r, Cc, device=dev).bfloat16()
    oc = torch.empty_like(xc)
    ms = timed(lambda: rt.check(l.dff_fov_warp_cl(xc.data_ptr(), alpha.data_ptr(), fov.data_ptr(), B, Cc, S, H // r, W // r, oc.data_ptr(), rt.BF16, 0, st())))
    rec("fov_warp_cl bf16 C=%d @1/%d, 8 x C4" % (Cc, r), ms, 2 * 2 * xc.numel(), "channels-last feature volume")
    del xc, oc
del x, out
# depth heads at the C2 shape, 16 stacks, fd as S scalars (the staged-input path) and tiled (the reference's tensor)
B, S, H, W = 16, 10, 384, 576
costs = [torch.randn(B, S, H // r, W // r, device=dev) * 8 for r in (8, 4, 2, 1)]
outs = [torch.empty(B, H, W, device=dev) for _ in range(4)]
cb = sum(c.numel() for c in costs) * 4 + 4 * 4 * B * H * W
f = l.dff_depth_heads4
f.restype = ctypes.c_int
f.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)] + [ctypes.c_int] * 4 + [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
cp = (ctypes.c_void_p * 4)(*[c.data_ptr() for c in costs])
op = (ctypes.c_void_p * 4)(*[o.data_ptr() for o in outs])
for name, fd in (("scalars", torch.linspace(0.28, 0.02, S, device=dev).view(1, S, 1, 1).expand(B, S, 1, 1).contiguous().expand(B, S, H, W)),
                 ("tiled", torch.linspace(0.28, 0.02, S, device=dev).view(1, S, 1, 1).expand(B, S, H, W).contiguous())):
    strides = (ctypes.c_int64 * 4)(*fd.stride())
    for fast in (1, 0):
        ms = timed(lambda: rt.check(f(cp, fd.data_ptr(), strides, B, S, H, W, op, fast, 0, st())))
        nb = cb + (4 * B * S if name == "scalars" else 4 * B * S * H * W)
        rec("depth_heads x4, fd %s, %s" % (name, "bf16-mode kernel (4 px/thread, SFU softplus)" if fast else "fp32-mode kernel"), ms, nb,
            "4 cost volumes + focus_dists in, 4 maps out; 16 DDFF stacks")
# uint8 staging -> pair-packed bf16 first-layer input
u8 = torch.randint(0, 256, (B, S, 383, 552, 3), dtype=torch.uint8, device=dev)
FS = torch.empty(B, 3, S, H, W, device=dev)
ms = timed(lambda: rt.check(l.dff_stage_u8(u8.data_ptr(), 383, 552, B, S, H, W, FS.data_ptr(), 0, st())))
rec("stage_u8 -> fp32 (B,3,S,H,W)", ms, u8.numel() + 4 * FS.numel(), "normalise + pad + transpose")
print(json.dumps(res, indent=1))
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)
