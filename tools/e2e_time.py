"""Ad-hoc device timing of the End-to-End forward (alignment + depth) on a simulator-shaped stack (BASELINE.json configs[3]):
   python tools/e2e_time.py [H] [W] [precision]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dffinthewild_b200.End_to_End import Network
from dffinthewild_b200 import synth

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) >= 3 else (512, 768)
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
torch.manual_seed(0)
net = Network()
net.load_state_dict(synth.synthetic_state(net.state_dict(), seed=2))
net.DFF_net.precision = prec
net.optical_flow_aggregation.precision = prec
net = net.cuda().eval()
FS, fd, fov = synth.focal_stack(1, 10, H, W).cuda(), synth.focus_dists(1, 10, H, W, "ddff", tiled=False).cuda(), synth.fovs(1, 10).cuda()
with torch.no_grad():
    for _ in range(10):
        net(FS, fd, fov)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    n = 20
    ta = tb = 0.0
    for _ in range(n):
        e[0].record(); w = net.optical_flow_aggregation(FS, fov); e[1].record(); net.DFF_net(w, fd); e[2].record()
        torch.cuda.synchronize()
        ta += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
V = 10 * H * W
print("E2E 1x10x3x%dx%d: alignment network %.2f ms [%.1f TFLOP/s, 56,504 FLOP/voxel], depth network (%s) %.2f ms -> %.1f stacks/s" % (
    H, W, ta / n, 56504.0 * V / (ta / n) / 1e9, prec, tb / n, 1000 * n / (ta + tb)))
