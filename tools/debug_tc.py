import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from dffinthewild_b200 import runtime as rt
torch.manual_seed(0)
def step(name, fn):
    try:
        r = fn(); torch.cuda.synchronize(); print("OK  ", name); return r
    except Exception as e:
        print("FAIL", name, "->", str(e)[:300]); return None
x = (torch.rand(1, 16, 2, 16, 32) * 2 - 1).bfloat16().float()
step("to_cl fp32", lambda: rt.to_channels_last(x.cuda(), 16, False))
step("to_cl bf16", lambda: rt.to_channels_last(x.cuda(), 16, True))
w = (torch.randn(16, 16, 3, 3, 3) * 0.05).bfloat16().float()
ref = F.conv3d(x.double(), w.double(), None, 1, 1)
step("ffma bf16 conv", lambda: rt.conv3d(x.cuda(), w.cuda(), bf16=True))
for cin, cout in ((16, 16), (32, 32), (64, 64), (8, 8), (8, 16), (16, 8), (128, 128)):
    x = (torch.rand(1, cin, 3, 16, 32) * 2 - 1).bfloat16().float()
    w = (torch.randn(cout, cin, 3, 3, 3) * (1.0 / (27 * cin)) ** 0.5).bfloat16().float()
    ref = F.conv3d(x.double(), w.double(), None, 1, 1)
    out = step("tc conv %d->%d" % (cin, cout), lambda: rt.conv3d(x.cuda(), w.cuda(), bf16=True, tensor_cores=int(os.environ.get("TCK", "1"))))
    if out is not None:
        err = (out.cpu().double() - ref).abs()
        print("     max err %.4g  (ref max %.3g)  mean err %.3g" % (err.max(), ref.abs().max(), err.mean()))
        if err.max() > 0.02 * ref.abs().max():
            o = out.cpu().double()
            # per-channel / per-slice error structure
            print("     err by out-channel:", [round(v, 3) for v in err.amax(dim=(0, 2, 3, 4)).tolist()][:16])
            print("     err by slice:", [round(v, 3) for v in err.amax(dim=(0, 1, 3, 4)).tolist()])
            print("     err by x (first 32):", [round(v, 2) for v in err.amax(dim=(0, 1, 2, 3)).tolist()][:32])
            print("     err by y:", [round(v, 2) for v in err.amax(dim=(0, 1, 2, 4)).tolist()])
