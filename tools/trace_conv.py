"""Pipeline trace / timing of ONE conv layer: python tools/trace_conv.py cin cout kd kh kw stride B S H W [transposed]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dffinthewild_b200 import runtime as rt
cin, cout, kd, kh, kw, stride, B, S, H, W = [int(a) for a in sys.argv[1:11]]
tr = len(sys.argv) > 11 and sys.argv[11] == "1"
x = torch.randn(B, cin, S, H, W, device="cuda")
w = torch.randn(*((cin, cout) if tr else (cout, cin)), kd, kh, kw, device="cuda") * 0.05
for _ in range(2):
    rt.conv3d(x, w, stride_hw=stride, transposed=tr, bf16=True, tensor_cores=1)
torch.cuda.synchronize()
print("done")
