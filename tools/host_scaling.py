"""dff_forward_host_u8: time per call as a function of the stacks per call (fixed cost of a synchronous call vs marginal cost)
   and of pieces of the pipeline switched off.   python tools/host_scaling.py"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from dffinthewild_b200 import runtime as rt, synth

dev = torch.device("cuda", 0)
net, sd = B.make_net("bf16")
net = net.to(dev).eval()
lib = rt.lib()
S, H, W = B.S, B.H, B.W
H0, W0 = B.VALID_HW
N = 192
hU8 = B.u8_stacks(N, 100).pin_memory()
hfd = synth.focus_dists(N, S, H, W, "ddff", tiled=False).pin_memory()
houts = [torch.empty((N, H, W), dtype=torch.float32).pin_memory() for _ in range(4)]
packed = rt.packed_weights(net.DFF_net, dev)
strides = (ctypes.c_int64 * 4)(S, 1, 0, 0)
mb = 64
dev_io = torch.empty(lib.dff_host_io_bytes_u8(mb, S, H0, W0, H, W, strides), dtype=torch.uint8, device=dev)
ws = torch.empty(lib.dff_workspace_bytes(mb, S, H, W, rt.BF16), dtype=torch.uint8, device=dev)
sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for n in (64, 128, 192):
    for with_out in (True, False):
        hp = (ctypes.c_void_p * 4)(*[(o.data_ptr() if with_out else None) for o in houts])
        f = lambda: rt.check(lib.dff_forward_host_u8(packed.data_ptr(), hU8.data_ptr(), H0, W0, hfd.data_ptr(), strides, n, mb, S, H, W, hp,
                                                     dev_io.data_ptr(), ws.data_ptr(), ws.numel(), rt.BF16, 0, sp))
        f(); f()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            f()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 5 * 1e3
        print("%3d stacks per call, D2H %s: %.2f ms per call = %.2f ms per 64 stacks = %.0f stacks/s" % (n, "on " if with_out else "off", ms, ms * 64 / n, n / ms * 1e3))
