import csv,sys,subprocess,io,re,collections
rep=sys.argv[1]; kid=int(sys.argv[2]) if len(sys.argv)>2 else 1; npart=float(sys.argv[3]) if len(sys.argv)>3 else 14155776
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
n=0;ops=collections.Counter();samp=collections.Counter();h=None
for r in rows:
    if r and r[0]=="Kernel Name": n+=1; continue
    if r and r[0]=="Address": h=r; continue
    if n==kid and h and len(r)>8:
        m=re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1])
        if not m: continue
        op=m.group(2).split('.')[0]
        ops[op]+=int(r[h.index("Instructions Executed")]); samp[op]+=int(r[h.index("# Samples")])
tot=sum(ops.values()); ts=sum(samp.values()); unit=npart/32
print("total",tot,"per 32 particles",round(tot/unit,1))
print("  ".join(f"{k}:{v/unit:.0f}({100*samp[k]/ts:.0f}%)" for k,v in ops.most_common(30)))
