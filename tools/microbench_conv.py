"""Kernel-only timing of single conv layers (run under ncu to read gpu__time_duration per launch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dffinthewild_b200 import runtime as rt
cases = [  # cin, cout, k, H, W  (B=4, S=10)
    (16, 16, (3, 3, 3), 192, 288), (16, 32, (3, 3, 3), 192, 288), (16, 64, (3, 3, 3), 192, 288),
    (32, 32, (3, 3, 3), 96, 144), (64, 32, (3, 3, 3), 96, 144), (16, 16, (1, 1, 1), 192, 288), (16, 16, (1, 3, 3), 192, 288),
    (64, 64, (1, 3, 3), 96, 144), (8, 8, (3, 3, 3), 384, 576),
    (8, 8, (1, 3, 3), 384, 576), (16, 8, (3, 3, 3), 384, 576), (32, 16, (3, 3, 3), 192, 288), (32, 32, (3, 3, 3), 192, 288),
]
mode = int(os.environ.get("MB_KERNEL", "1"))
for cin, cout, k, H, W in cases:
    x = torch.randn(4, cin, 10, H, W, device="cuda")
    w = torch.randn(cout, cin, *k, device="cuda") * 0.05
    rt.conv3d(x, w, bf16=True, tensor_cores=mode)
torch.cuda.synchronize()
print("done")
