"""Join an ncu launch list of a training step with the DFF_B200_WGRAD_LOG lines of the same run:
   python tools/wgrad_by_layer.py launches.csv stderr.log"""
import csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
w = [float(x['Metric Value'].replace(',', '')) / 1000 for x in rows if 'wgrad_tma' in x['Kernel Name'] or 'wgrad_mma' in x['Kernel Name']]
log = [l.strip() for l in open(sys.argv[2]) if l.startswith('wgrad_')]
n = len(w); log = log[-n:]
print("%d weight-gradient launches, %.1f us in total" % (n, sum(w)))
for t, l in zip(w, log):
    d = dict(re.findall(r'(\w+)=([\d+]+)', l))
    cin = sum(int(x) for x in d['Cin'].split('+')); cout = int(d['Cout']); taps = int(d['taps'])
    pos = int(d['B']) * int(d['S']) * int(d['OHt']) * int(d['OWt'])
    gf = 2 * cin * cout * taps * pos / 1e9
    print("%7.1f us %6.2f GFLOP %6.1f TFLOP/s  %s" % (t, gf, gf / t * 1e3, l))
