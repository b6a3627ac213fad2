"""Experiment: 64 DDFF stacks per step as ONE dff_forward_u8 call vs k concurrent calls of 64/k stacks on k streams (tail filling).
   python tools/two_stream.py"""
import ctypes, os, sys
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
n = 64
U8 = B.u8_stacks(n, 100).to(dev)
fd = synth.focus_dists(n, S, H, W, "ddff", tiled=False).to(dev)
outs = [torch.empty((n, H, W), dtype=torch.float32, device=dev) for _ in range(4)]
packed = rt.packed_weights(net.DFF_net, dev)
strides = (ctypes.c_int64 * 4)(S, 1, 0, 0)
main = torch.cuda.current_stream(dev)
ref = None
for k in (1, 2, 4):
    mb = n // k
    streams = [torch.cuda.Stream(device=dev) for _ in range(k)]
    wss = [torch.empty(lib.dff_workspace_bytes(mb, S, H, W, rt.BF16), dtype=torch.uint8, device=dev) for _ in range(k)]

    def step():
        for j, st in enumerate(streams):
            st.wait_stream(main)
            i = j * mb
            op = (ctypes.c_void_p * 4)(*[o[i:i + mb].data_ptr() for o in outs])
            rt.check(lib.dff_forward_u8(packed.data_ptr(), U8[i:i + mb].data_ptr(), H0, W0, fd[i:i + mb].data_ptr(), strides, mb, S, H, W,
                                        op, None, wss[j].data_ptr(), wss[j].numel(), rt.BF16, 0, ctypes.c_void_p(st.cuda_stream)))
        for st in streams:
            main.wait_stream(st)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(20):
        step()
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    if ref is None:
        ref = [o.clone() for o in outs]
    same = all(torch.equal(a, b) for a, b in zip(ref, outs))
    print("%d stream(s) x %d stacks: %.2f ms per 64 stacks = %.0f stacks/s (bit-identical to the single call: %s)" % (k, mb, ms, n / ms * 1e3, same))
    del wss
