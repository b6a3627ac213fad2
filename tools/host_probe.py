"""Why is the host pipeline's steady state ~5 % slower than one device-resident 64-stack call?  (a) the chunk schedule itself on
   device-resident inputs, (b) a 64-stack call with unrelated H2D / D2H copies in flight.   python tools/host_probe.py"""
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
hU8 = B.u8_stacks(n, 100).pin_memory()
U8 = hU8.to(dev)
fd = synth.focus_dists(n, S, H, W, "ddff", tiled=False).to(dev)
outs = [torch.empty((n, H, W), dtype=torch.float32, device=dev) for _ in range(4)]
houts = [torch.empty((n, H, W), dtype=torch.float32).pin_memory() for _ in range(4)]
packed = rt.packed_weights(net.DFF_net, dev)
strides = (ctypes.c_int64 * 4)(S, 1, 0, 0)
main = torch.cuda.current_stream(dev)
ws = torch.empty(lib.dff_workspace_bytes(64, S, H, W, rt.BF16), dtype=torch.uint8, device=dev)
half = (ws.numel() // 2) & ~255
s2, cp1, cp2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)


def fwd(i, m, stream, wsp, wsz):
    op = (ctypes.c_void_p * 4)(*[o[i:i + m].data_ptr() for o in outs])
    rt.check(lib.dff_forward_u8(packed.data_ptr(), U8[i:i + m].data_ptr(), H0, W0, fd[i:i + m].data_ptr(), strides, m, S, H, W, op, None,
                                wsp, wsz, rt.BF16, 0, ctypes.c_void_p(stream.cuda_stream)))


def timed(step, label):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(10):
        step()
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%-70s %.2f ms per 64 stacks = %.0f stacks/s" % (label, ms, n / ms * 1e3))


timed(lambda: fwd(0, 64, main, ws.data_ptr(), ws.numel()), "one call, 64 stacks")


def sched(sizes):
    def step():
        s2.wait_stream(main)
        i = 0
        for k, m in enumerate(sizes):
            fwd(i, m, s2 if k & 1 else main, ws.data_ptr() + (half if k & 1 else 0), half)
            i += m
        main.wait_stream(s2)
    return step


for sizes in ([2, 6, 24, 24, 6, 2], [32, 32], [8, 24, 24, 8]):
    timed(sched(sizes), "device-resident, two streams, chunks %s" % sizes)


def with_copies():
    cp1.wait_stream(main); cp2.wait_stream(main)
    with torch.cuda.stream(cp1):
        U8.copy_(hU8, non_blocking=True)
    fwd(0, 64, main, ws.data_ptr(), ws.numel())
    with torch.cuda.stream(cp2):
        for h, o in zip(houts, outs):
            h.copy_(o, non_blocking=True)
    main.wait_stream(cp1); main.wait_stream(cp2)


timed(with_copies, "one call, 64 stacks, 406 MB H2D + 226 MB D2H in flight")
