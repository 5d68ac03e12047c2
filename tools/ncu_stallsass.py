import csv,sys,subprocess,io
rep=sys.argv[1]; kid=int(sys.argv[2]); col=sys.argv[3]; top=int(sys.argv[4]) if len(sys.argv)>4 else 25
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
n=0;h=None;body=[]
for r in rows:
    if r and r[0]=="Kernel Name": n+=1; continue
    if r and r[0]=="Address": h=r; continue
    if n==kid and h and len(r)>40: body.append(r)
ci=h.index(col); si=h.index("# Samples")
tot=sum(int(r[ci]) for r in body); ts=sum(int(r[si]) for r in body)
print(col,"total",tot,"of",ts,"samples")
idx=sorted(range(len(body)), key=lambda i:-int(body[i][ci]))[:top]
for i in idx:
    prev=body[i-1][1].strip()[:60] if i>0 else ""
    print(f"{i:5d} {100*int(body[i][ci])/ts:5.2f}%  {body[i][1].strip()[:70]:70s} | prev: {prev}")
