import csv,sys,subprocess,io
rep=sys.argv[1]; kid=int(sys.argv[2]) if len(sys.argv)>2 else 1; top=int(sys.argv[3]) if len(sys.argv)>3 else 30
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
out=[];n=0
for r in rows:
    if r and r[0]=="Function Name": n+=1
    if n==kid and len(r)>8 and r[0] not in("Line No",""):
        try: out.append((int(r[0]), r[1], int(r[7]), int(r[6])))
        except: pass
ts=sum(o[3] for o in out); tot=sum(o[2] for o in out)
for ln,src,c,s in sorted(out,key=lambda o:-o[3])[:top]:
    print(f"{ln:5d} samp {100*s/ts:5.1f}%  instr {100*c/tot:5.1f}%  {src[:100]}")
