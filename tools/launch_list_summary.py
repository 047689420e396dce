import csv, sys, collections
def load(path):
    rows=[r for r in csv.reader(open(path)) if len(r)>10]
    hdr=rows[0]
    iK,iM,iV,iID,iU=hdr.index("Kernel Name"),hdr.index("Metric Name"),hdr.index("Metric Value"),hdr.index("ID"),hdr.index("Metric Unit")
    L=collections.OrderedDict()
    for r in rows[1:]:
        v=float(r[iV].replace(",",""))
        if r[iM]=="gpu__time_duration.sum": v*={"ns":1e-3,"us":1.0,"ms":1e3,"s":1e6}.get(r[iU],1e-3)
        k=r[iK].split("(")[0].replace("void ","").replace("nrcu::","").strip()
        L.setdefault(r[iID],{"kernel":k})[r[iM]]=v
    return list(L.values())
for path in sys.argv[1:]:
    L=load(path)
    agg=collections.OrderedDict()
    for l in L:
        a=agg.setdefault(l["kernel"],[0,0.0,0.0,0.0,0.0])
        a[0]+=1; a[1]+=l["gpu__time_duration.sum"]; a[2]+=l["smsp__inst_executed.sum"]
        a[3]+=l["smsp__inst_executed.sum"]*l["smsp__thread_inst_executed_per_inst_executed.ratio"]
        a[4]+=l["gpu__time_duration.sum"]*l["smsp__issue_active.avg.pct_of_peak_sustained_active"]
    print(path, len(L))
    T=sum(a[1] for a in agg.values()); I=sum(a[2] for a in agg.values())
    for k,a in agg.items():
        print(f"  {k:28s} n={a[0]:4d} time {a[1]:10.0f} us ({a[1]/T*100:5.1f}%) winst {a[2]/1e6:9.1f} M ({a[2]/I*100:5.1f}%) lanes {a[3]/max(a[2],1):5.2f} issue% {a[4]/max(a[1],1e-9):5.1f}")
    print(f"  total time {T:.0f} us, winst {I/1e6:.1f} M")
