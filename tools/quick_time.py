"""Ad-hoc device timing of the whole forward (not the bench): python tools/quick_time.py [B] [S] [H] [W] [precision]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dffinthewild_b200.Depth_Estimation_Network import Network
from dffinthewild_b200 import synth

B, S, H, W = [int(a) for a in sys.argv[1:5]] if len(sys.argv) >= 5 else (2, 10, 384, 576)
prec = sys.argv[5] if len(sys.argv) > 5 else "fp32"
torch.manual_seed(0)
net = Network()
net.load_state_dict(synth.synthetic_state(net.state_dict(), seed=1))
net.DFF_net.precision = prec
net = net.cuda().eval()
FS, fd = synth.focal_stack(B, S, H, W).cuda(), synth.focus_dists(B, S, H, W).cuda()
with torch.no_grad():
    for _ in range(2):
        net(FS, fd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 3
    for _ in range(n):
        net(FS, fd)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
V = S * H * W
print("forward %s B=%d S=%d %dx%d: %.2f ms/batch, %.2f ms/stack, %.1f stacks/s, %.2f TFLOP/s (93,563 FLOP/voxel)" % (
    prec, B, S, H, W, ms, ms / B, 1000 * B / ms, 93563.0 * V * B / ms / 1e9))
