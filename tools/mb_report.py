import csv, sys
cases=[(16,16,27,192,288),(16,32,27,192,288),(16,64,27,192,288),(32,32,27,96,144),(64,32,27,96,144),(16,16,1,192,288),(16,16,9,192,288),(64,64,9,96,144),(8,8,27,384,576)]
for f in sys.argv[1:]:
    rows=list(csv.reader(open(f)))
    hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]; h=rows[hdr]
    vi=h.index('Metric Value'); gi=h.index('Grid Size')
    print(f)
    for r,(ci,co,t,H,W) in zip(rows[hdr+1:],cases):
        ns=float(r[vi].replace(',','')); tiles=4*10*(H//16)*(W//8); mm=t*ci//16 if ci>8 else (15 if t==27 else (t+1)//2)
        print("%3d->%3d taps %2d  %8.1f us grid %-12s clk/MMA/SM %.0f  TFLOP/s %.0f"%(ci,co,t,ns/1e3,r[gi],ns*1.9*148/tiles/mm, 2*4*10*H*W*ci*co*t/ns/1e3))
