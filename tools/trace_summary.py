"""Per-launch role times of the slab kernel from a whole-forward trace (tools/one_forward.py with the -DDFF_SLAB_TRACE build):
   python tools/trace_summary.py trace.txt launches.csv B S H W
cadence = clocks between consecutive slices of CTA 0; prod = empty-wait -> loads issued (cp.async producers only; TMA: n/a);
mma = accumulator free -> MMAs issued; epi = accumulator full -> epilogue done."""
import re, statistics as st, ctypes, sys, os, csv
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dffinthewild_b200 import runtime as rt
tr, lc, B, S, H, W = sys.argv[1], sys.argv[2], *[int(a) for a in sys.argv[3:7]]
lines = open(tr).read().split('\n')
idx = [i for i, l in enumerate(lines) if l.startswith('trace')]
rows = list(csv.reader(open(lc)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]; h = rows[hdr]; data = [r for r in rows[hdr + 1:] if len(r) == len(h)]
ki = h.index('Kernel Name'); vi = h.index('Metric Value')
first = next(i for i, r in enumerate(data) if 'to_cl_pair' in r[ki]); data = data[first:first + 300]
l = rt.lib(); N = 256
ms = (ctypes.c_float * N)(); fl = (ctypes.c_double * N)(); by = (ctypes.c_double * N)(); la = (ctypes.c_int * N)(); nm = ctypes.create_string_buffer(N * 64); n = ctypes.c_int(0)
l.dff_forward_profiled(None, None, None, None, B, S, H, W, None, None, 0, 1, 0, None, N, ms, fl, by, la, nm, ctypes.byref(n))
names = []
for k in range(n.value): names += [nm.raw[k * 64:(k + 1) * 64].split(b"\0")[0].decode()] * la[k]
slab = [(names[i], float(data[i][vi].replace(',', '')) / 1e3) for i in range(len(names)) if 'conv_slab' in data[i][ki]]
print("%-46s %7s | %5s %3s %3s %2s | %7s %7s %7s %7s" % ("layer", "us", "N", "ops", "NP", "oc", "cadence", "prod", "mma", "epi"))
for j, i in enumerate(idx):
    m = re.search(r'N=(\d+) nops=(\d+) NP=(\d+) occ=(\d+) grid=(\d+) wstream=(\d+)', lines[i])
    rd = []
    for l2 in lines[i + 1:i + 26]:
        p = l2.split('|')[0].split()
        if len(p) >= 7 and p[0] != '-1': rd.append([int(x) for x in p[:7]])
    if len(rd) < 4: continue
    cad = st.median([rd[k + 1][0] - rd[k][0] for k in range(len(rd) - 1)])
    prod = st.median([r[6] - r[5] for r in rd]) if rd[0][5] >= 0 and rd[1][6] > 0 else -1
    mma = st.median([r[2] - r[1] for r in rd]); epi = st.median([r[4] - r[3] for r in rd])
    nme, us = slab[j] if j < len(slab) else ('?', 0)
    print("%-46s %7.1f | %5s %3s %3s %2s | %7d %7d %7d %7d %s" % (nme[:46], us, m.group(1), m.group(2), m.group(3), m.group(4), cad, prod, mma, epi, 'WS' if m.group(6) == '1' else ''))
