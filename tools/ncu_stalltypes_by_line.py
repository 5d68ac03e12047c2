"""Stall reasons per CUDA source line (top lines by samples). usage: rep kernel_index [top]"""
import csv,sys,subprocess,io,collections
rep=sys.argv[1]; kid=int(sys.argv[2]); top=int(sys.argv[3]) if len(sys.argv)>3 else 25
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
n=0;h=None;out=[]
for r in rows:
    if r and r[0]=="Function Name": n+=1
    if r and r[0]=="Line No": h=r; continue
    if n==kid and h and len(r)>40 and r[0] not in("Line No",""):
        try: int(r[0]); int(r[h.index("# Samples")])
        except: continue
        out.append(r)
names=[c for c in h if c.startswith("stall_") and "Not Issued" not in c]
si=h.index("# Samples"); ts=sum(int(r[si]) for r in out)
tot=collections.Counter()
for r in out:
    for c in names: tot[c]+=int(r[h.index(c)])
print("all lines:", "  ".join(f"{c[6:]}:{100*v/ts:.1f}%" for c,v in tot.most_common(9)))
for r in sorted(out,key=lambda r:-int(r[si]))[:top]:
    d={c:int(r[h.index(c)]) for c in names}
    s=int(r[si])
    print(f"{int(r[0]):5d} {100*s/ts:5.1f}%  "+"  ".join(f"{c[6:]}:{100*v/max(s,1):.0f}%" for c,v in sorted(d.items(),key=lambda kv:-kv[1])[:4])+"   | "+r[1].strip()[:60])
