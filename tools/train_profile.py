"""One graph-replayed training step of the C3 shape for the profiler (warm-up and capture outside the cudaProfilerStart/Stop window):
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/train_profile.py [precision] [graph 0|1]
   then python tools/launch_summary.py launches.csv"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dffinthewild_b200 import synth, train_step as TS
from dffinthewild_b200.Depth_Estimation_Network import Network

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
graph = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
B, S, H, W = (int(sys.argv[3]) if len(sys.argv) > 3 else 4), 5, 256, 256
torch.manual_seed(0)
net = Network()
net.load_state_dict(synth.synthetic_state(net.state_dict(), seed=1))
net.DFF_net.precision = prec
net = net.cuda().train()
FS, fd = synth.focal_stack(B, S, H, W).cuda(), synth.focus_dists(B, S, H, W, "defocus", tiled=False).cuda()
gt, mask = synth.gt_and_mask(B, H, W)
gt, mask = gt.cuda(), mask.cuda()
st = TS.TrainStep(net, lr=1e-4, use_graph=graph)
for _ in range(4):
    info = st.step(FS, fd, gt, mask)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    st.step(FS, fd, gt, mask)
e1.record()
torch.cuda.synchronize()
print("train step %s graph=%s B=%d: %.2f ms/step = %.0f stacks/s (graphed: %s)" % (prec, graph, B, e0.elapsed_time(e1) / 5, 5000.0 * B / e0.elapsed_time(e1), info["graphed"]))
torch.cuda.profiler.start()
st.step(FS, fd, gt, mask)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
