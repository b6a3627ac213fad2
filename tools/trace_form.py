"""Pipeline trace of one layer in the form dff_forward runs it (plan 4): DFF_B200_LIB=tools/libdff_trace.so DFF_SLAB_TRACE=1 python tools/trace_form.py <case> [B]
   cases: conv6_last (dres4.conv6: deconv 16->8 + BN + skip after BN + fused classifier, nothing else stored), deconv3 (deconv 16->8 + BN),
          conv0 (dres4.conv0: 8+8 -> 8 3x3x3 + BN + ReLU), conv5 (dres4.conv5: deconv 16->16?), first (1x9x9 dil 2, 3 -> 8)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dffinthewild_b200 import runtime as rt
case = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
S, H, W = 10, 384, 576
g = torch.Generator().manual_seed(0)
r = lambda *s: (torch.rand(*s, generator=g) * 2 - 1).cuda()
if case == "conv6_last":
    x, w = r(B, 16, S, H // 2, W // 2), r(16, 8, 3, 3, 3) * 0.1
    f = lambda: rt.conv3d_forward_plan(x, w, 2, 1, True, scale=r(8) + 1.5, shift=r(8), res_post=r(B, 8, S, H, W), proj_w=r(8), proj_on_aux=False, skip_out=True)
elif case == "deconv3":
    x, w = r(B, 16, S, H // 2, W // 2), r(16, 8, 3, 3, 3) * 0.1
    f = lambda: rt.conv3d_forward_plan(x, w, 2, 1, True, scale=r(8) + 1.5, shift=r(8))
elif case == "conv0":
    x, x2, w = r(B, 8, S, H, W), r(B, 8, S, H, W), r(8, 16, 3, 3, 3) * 0.1
    f = lambda: rt.conv3d_forward_plan(x, w, 1, 1, False, scale=r(8) + 1.5, shift=r(8), relu=True, x2=x2)
elif case == "conv5":
    x, w = r(B, 32, S, H // 4, W // 4), r(32, 16, 3, 3, 3) * 0.1
    f = lambda: rt.conv3d_forward_plan(x, w, 2, 1, True, scale=r(16) + 1.5, shift=r(16), res_pre=r(B, 16, S, H // 2, W // 2), relu=True)
elif case == "srd0":   # FM_measure...Focus_Measure.conv.0: 8 -> 8, 1x3x3, BN + ReLU (x-folded G = 4)
    x, w = r(B, 8, S, H, W), r(8, 8, 1, 3, 3) * 0.2
    f = lambda: rt.conv3d_forward_plan(x, w, 1, 1, False, scale=r(8) + 1.5, shift=r(8), relu=True)
elif case == "srd2":   # ...conv.2: + residual before the ReLU
    x, w = r(B, 8, S, H, W), r(8, 8, 1, 3, 3) * 0.2
    f = lambda: rt.conv3d_forward_plan(x, w, 1, 1, False, scale=r(8) + 1.5, shift=r(8), res_pre=r(B, 8, S, H, W), relu=True)
elif case == "first":
    x, w = r(B, 3, S, H, W), r(8, 3, 1, 9, 9) * 0.1
    f = lambda: rt.conv3d_forward_plan(x, w, 1, 2, False, scale=r(8) + 1.5, shift=r(8), relu=True)
f()
torch.cuda.synchronize()
print("done")
