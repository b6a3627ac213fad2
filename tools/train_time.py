"""Ad-hoc device timing of one training step (fwd + Defocus loss + bwd + Adam): python tools/train_time.py [B] [S] [H] [W]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dffinthewild_b200.Depth_Estimation_Network import Network
from dffinthewild_b200 import synth

B, S, H, W = [int(a) for a in sys.argv[1:5]] if len(sys.argv) >= 5 else (4, 5, 256, 256)
prec = sys.argv[5] if len(sys.argv) > 5 else "fp32"
torch.manual_seed(0)
net = Network()
net.load_state_dict(synth.synthetic_state(net.state_dict(), seed=1))
net.DFF_net.precision = prec
net = net.cuda().train()
opt = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.9, 0.99))
FS, fd = synth.focal_stack(B, S, H, W).cuda(), synth.focus_dists(B, S, H, W, "defocus").cuda()
gt, mask = synth.gt_and_mask(B, H, W)
gt, mask = gt.cuda(), mask.cuda()
crit = torch.nn.MSELoss()

def step():
    o = net(FS, fd)
    opt.zero_grad()
    loss = 0.5 * crit(o[1][mask], gt[mask]) + 0.7 * crit(o[2][mask], gt[mask]) + crit(o[3][mask], gt[mask]) + 0.3 * crit(o[0][mask], gt[mask])
    loss.backward()
    opt.step()
    return loss

for _ in range(2):
    step()
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
e[0].record(); o = net(FS, fd); e[1].record()
loss = 0.5 * crit(o[1][mask], gt[mask]) + 0.7 * crit(o[2][mask], gt[mask]) + crit(o[3][mask], gt[mask]) + 0.3 * crit(o[0][mask], gt[mask])
opt.zero_grad(); loss.backward(); e[2].record(); opt.step(); e[3].record()
torch.cuda.synchronize()
V = S * H * W
print("train step " + prec + " B=%d S=%d %dx%d: fwd %.1f ms, loss+bwd %.1f ms, adam %.1f ms -> %.2f stacks/s, %.2f TFLOP/s (276,801 FLOP/voxel)" % (
    B, S, H, W, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), 1000 * B / e[0].elapsed_time(e[3]),
    276801.0 * V * B / e[0].elapsed_time(e[3]) / 1e9))
