import csv, sys
cases=[(16,16,27,192,288),(16,32,27,192,288),(16,64,27,192,288),(32,32,27,96,144),(64,32,27,96,144),(16,16,1,192,288),(16,16,9,192,288),(64,64,9,96,144),(8,8,27,384,576),
       (8,8,9,384,576),(16,8,27,384,576),(32,16,27,192,288),(32,32,27,192,288)]
for f in sys.argv[1:]:
    rows=list(csv.reader(open(f)))
    hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]; h=rows[hdr]
    vi=h.index('Metric Value'); gi=h.index('Grid Size'); ki=h.index('Kernel Name')
    print(f)
    for r,(ci,co,t,H,W) in zip(rows[hdr+1:],cases):
        ns=float(r[vi].replace(',','')); tiles=4*10*H*W/128; mm=t*ci//16 if ci>8 else (15 if t==27 else (t+1)//2)
        hbm = 4*10*H*W*(ci+co)*2/6.45e3   # ns at 6.45 TB/s
        print("%3d->%3d taps %2d %dx%d %-10s %8.1f us grid %-12s clk/128px %5.0f  TFLOP/s %4.0f  x HBM-roofline %.1f"%(ci,co,t,H,W,r[ki][5:14],ns/1e3,r[gi],ns*1.9*148/tiles, 2*4*10*H*W*ci*co*t/ns/1e3, ns/hbm))
