"""Top stall sites of one kernel from `ncu --page source --csv` (SASS view): python tools/ncu_hot.py src.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) >= len(h) - 2]
si, ns, ie = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
stalls = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
tot = sum(int(r[ns] or 0) for r in data)
print("total samples", tot, " instructions", sum(int(r[ie] or 0) for r in data))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
order = sorted(range(len(data)), key=lambda k: -int(data[k][ns] or 0))[:n]
for k in sorted(order):
    r = data[k]
    top = sorted(((int(r[i] or 0), c) for i, c in stalls), reverse=True)[:2]
    print("%5d %5.1f%% exec %9s  %-70s %s" % (k, 100.0 * int(r[ns] or 0) / tot, r[ie], r[si].strip()[:70], " ".join("%s=%d" % (c[6:], v) for v, c in top if v)))
