"""Join an ncu launch list (one forward chunk) with the library's operator plan: per-layer time, clk/K-step, TFLOP/s.
   python tools/launch_by_layer.py launches.csv B S H W mode(0 fp32,1 bf16)"""
import csv, sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dffinthewild_b200 import runtime as rt
f, B, S, H, W, mode = sys.argv[1], *[int(a) for a in sys.argv[2:7]]
rows = list(csv.reader(open(f)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]; data = [r for r in rows[hdr + 1:] if len(r) == len(h)]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
first = next(i for i, r in enumerate(data) if ('to_cl_kernel' in r[ki] or 'to_cl_pair_kernel' in r[ki]))  # align to the start of one forward chunk
data = data[first:]
times = [float(r[vi].replace(',', '')) for r in data]
l = rt.lib(); N = 256
ms = (ctypes.c_float * N)(); fl = (ctypes.c_double * N)(); by = (ctypes.c_double * N)(); la = (ctypes.c_int * N)()
nm = ctypes.create_string_buffer(N * 64); n = ctypes.c_int(0)
l.dff_forward_profiled(None, None, None, None, B, S, H, W, None, None, 0, mode, 0, None, N, ms, fl, by, la, nm, ctypes.byref(n))
i = 0; tot = sum(times[:sum(la[k] for k in range(n.value))])
print("total %.3f ms for %d stacks" % (tot / 1e6, B))
for k in range(n.value):
    t = sum(times[i:i + la[k]]); i += la[k]
    name = nm.raw[k * 64:(k + 1) * 64].split(b"\0")[0].decode()
    print("%-52s %8.1f us %5.1f%%  %7.2f TFLOP/s  %7.1f GB/s" % (name, t / 1e3, 100 * t / tot, fl[k] / t / 1e3, by[k] / t))
