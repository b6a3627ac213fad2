"""Per-operator table of one forward (CUDA events around every operator, dff_forward_profiled): python tools/by_op.py [B] [precision] > table
Two tables (e.g. with and without an A/B environment knob) are compared with: python tools/by_op.py --diff a.txt b.txt"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1 and sys.argv[1] == "--diff":
    def load(f):
        d, order = {}, []
        for l in open(f):
            p = l.split()
            if len(p) >= 3 and p[-1] == "us":
                n = p[0]
                k = 1
                while n in d:
                    k += 1
                    n = "%s#%d" % (p[0], k)
                d[n] = float(p[-2]); order.append(n)
        return d, order
    a, order = load(sys.argv[2]); b, _ = load(sys.argv[3])
    ta = tb = 0.0
    for n in order:
        if n in b:
            ta += a[n]; tb += b[n]
            flag = "" if abs(b[n] - a[n]) < 0.02 * a[n] else ("  <<< faster" if b[n] < a[n] else "  >>> SLOWER")
            print("%-56s %9.1f %9.1f  %+7.1f%s" % (n, a[n], b[n], b[n] - a[n], flag))
    print("%-56s %9.1f %9.1f  %+7.1f" % ("TOTAL", ta, tb, tb - ta))
    sys.exit(0)

import torch
from dffinthewild_b200 import runtime as rt, synth
from dffinthewild_b200.Depth_Estimation_Network import Network

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
S, H, W = 10, 384, 576
if len(sys.argv) > 5:
    S, H, W = int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
torch.manual_seed(0)
net = Network()
net.load_state_dict(synth.synthetic_state(net.state_dict(), seed=1))
net.DFF_net.precision = prec
net = net.cuda().eval()
dev = torch.device("cuda", 0)
lib = rt.lib()
mode = rt.BF16 if prec == "bf16" else rt.FP32
FS = synth.focal_stack(B, S, H, W).cuda()
fdt = synth.focus_dists(B, S, H, W).cuda().expand(B, S, H, W).contiguous()
tstr = (ctypes.c_int64 * 4)(*fdt.stride())
outs = [torch.empty((B, H, W), dtype=torch.float32, device=dev) for _ in range(4)]
packed = rt.packed_weights(net.DFF_net, dev)
ws = torch.empty(lib.dff_workspace_bytes(B, S, H, W, mode), dtype=torch.uint8, device=dev)
sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
optr = (ctypes.c_void_p * 4)(*[o.data_ptr() for o in outs])
NOPS = 256
op_ms, op_fl, op_by = (ctypes.c_float * NOPS)(), (ctypes.c_double * NOPS)(), (ctypes.c_double * NOPS)()
op_la, op_nm, n_ops = (ctypes.c_int * NOPS)(), ctypes.create_string_buffer(NOPS * 64), ctypes.c_int(0)
REP = 4
acc = None
for j in range(REP + 1):
    rt.check(lib.dff_forward_profiled(packed.data_ptr(), FS.data_ptr(), fdt.data_ptr(), tstr, B, S, H, W, optr, ws.data_ptr(), ws.numel(),
                                      mode, 0, sp, NOPS, op_ms, op_fl, op_by, op_la, op_nm, ctypes.byref(n_ops)))
    if j == 0:
        continue   # warm-up
    if acc is None:
        acc = [0.0] * n_ops.value
    for k in range(n_ops.value):
        acc[k] += op_ms[k] / REP
tot = sum(acc)
print("total %.3f ms for %d stacks (%s)" % (tot, B, prec))
for k in range(n_ops.value):
    name = op_nm.raw[k * 64:(k + 1) * 64].split(b"\0")[0].decode().replace(" ", "_")
    t = acc[k]
    print("%-56s %6.1f TFLOP/s %7.1f GB/s %5.1f%% %9.1f us" % (name, op_fl[k] / (t / 1e3) / 1e12, op_by[k] / (t / 1e3) / 1e9, 100 * t / tot, t * 1e3))
