"""Per CUDA source line: warp instructions executed per 32 particles and stall-sample share, sorted by instructions.
usage: python tools/ncu_lineinst.py rep kernel_index nparticles [top]"""
import csv,sys,subprocess,io
rep=sys.argv[1]; kid=int(sys.argv[2]); npart=float(sys.argv[3]); top=int(sys.argv[4]) if len(sys.argv)>4 else 60
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
out=[];n=0
for r in rows:
    if r and r[0]=="Function Name": n+=1
    if n==kid and len(r)>8 and r[0] not in("Line No",""):
        try: out.append((int(r[0]), r[1], int(r[7]), int(r[6])))
        except: pass
ts=sum(o[3] for o in out); tot=sum(o[2] for o in out); unit=npart/32
print("total per 32 particles %.1f"%(tot/unit))
for ln,src,c,s in sorted(out,key=lambda o:-o[2])[:top]:
    print(f"{ln:5d} inst/32p {c/unit:7.1f}  samp {100*s/ts:5.1f}%  {src.strip()[:110]}")
