"""One profiled forward for ncu: warm-up forwards run outside the cudaProfilerStart/Stop window.
   ncu --profile-from-start off --set full ... python tools/one_forward.py [B] [S] [H] [W] [precision]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dffinthewild_b200.Depth_Estimation_Network import Network
from dffinthewild_b200 import synth

B, S, H, W = [int(a) for a in sys.argv[1:5]] if len(sys.argv) >= 5 else (16, 10, 384, 576)
prec = sys.argv[5] if len(sys.argv) > 5 else "bf16"
torch.manual_seed(0)
net = Network()
net.load_state_dict(synth.synthetic_state(net.state_dict(), seed=1))
net.DFF_net.precision = prec
net = net.cuda().eval()
FS, fd = synth.focal_stack(B, S, H, W, valid_hw=(H - 1, W - 24)).cuda(), synth.focus_dists(B, S, H, W, "ddff").cuda()
with torch.no_grad():
    for _ in range(2):
        net(FS, fd)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    net(FS, fd)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
