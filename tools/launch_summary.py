"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py file.csv [n_first]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]; data = [r for r in rows[hdr + 1:] if len(r) == len(h)]
ki, vi, gi, bi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size'), h.index('Block Size')
n = int(sys.argv[2]) if len(sys.argv) > 2 else len(data)
t = [(r[ki], float(r[vi].replace(',', '')), r[gi], r[bi]) for r in data][:n]
tot = sum(x[1] for x in t)
print("launches %d  total %.3f ms" % (len(t), tot / 1e6))
agg = collections.defaultdict(lambda: [0, 0.0])
for x in t:
    k = x[0][:70]
    agg[k][0] += 1; agg[k][1] += x[1]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s n=%3d  %.3f ms  %.1f%%" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot))
print("-- per launch (>=1.5%)")
for i, x in enumerate(t):
    if x[1] / tot >= 0.015:
        print("%3d %-40s %8.3f ms %5.1f%% grid %s" % (i, x[0][:40], x[1] / 1e6, 100 * x[1] / tot, x[2]))
