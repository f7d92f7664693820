import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr=None; cur=None
agg=collections.OrderedDict()
for r in rows:
    if len(r)==2 and r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if len(r)>10 and r[0]=="Line No": hdr={h:i for i,h in enumerate(r)}; continue
    if hdr is None or len(r)<10 or not r[0].isdigit(): continue
    key=(cur,int(r[0]))
    f=lambda v: float(v) if v not in ("","-") else 0.0; inst=f(r[hdr["Instructions Executed"]]); tinst=f(r[hdr["Thread Instructions Executed"]]); smp=f(r[hdr["# Samples"]])
    a=agg.setdefault(key,[0,0,0]); a[0]+=inst; a[1]+=tinst; a[2]+=smp
tot=sum(v[0] for v in agg.values()); tots=sum(v[2] for v in agg.values())
ranges=eval(sys.argv[2])  # list of (name, file, lo, hi)
for name,f,lo,hi in ranges:
    i=sum(v[0] for k,v in agg.items() if k[0]==f and lo<=k[1]<=hi); t=sum(v[1] for k,v in agg.items() if k[0]==f and lo<=k[1]<=hi); s=sum(v[2] for k,v in agg.items() if k[0]==f and lo<=k[1]<=hi)
    print(f"{name:28s} inst {100*i/tot:5.1f}%  samples {100*s/tots:5.1f}%  thr/inst {t/max(i,1):5.1f}")
