"""Shared-memory wavefronts per SASS opcode from an ncu report: python tools/ncu_smem_by_opcode.py report.ncu-rep [launch_index]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks = txt.split('"Kernel Name"')
lines = blocks[1 + which].splitlines()
hi = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = csv.reader(io.StringIO("\n".join(lines[hi:])))
hdr = next(rd); col = {n: i for i, n in enumerate(hdr)}
iw, ii, ix, isrc = col["L1 Wavefronts Shared"], col["L1 Wavefronts Shared Ideal"], col["Instructions Executed"], col["Source"]
agg = collections.defaultdict(lambda: [0, 0, 0])
for row in rd:
    if len(row) < len(hdr): continue
    toks = row[isrc].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith('@') else toks[0]
    try: w = float(row[iw]); idl = float(row[ii]); n = float(row[ix])
    except ValueError: continue
    if w > 0:
        a = agg[op]; a[0] += w; a[1] += idl; a[2] += n
tot = sum(a[0] for a in agg.values())
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-22s wavefronts %5.1f%% %.3g  ideal %.3g  inst %.3g  wf/inst %.2f  ideal/inst %.2f" % (op, 100*a[0]/tot, a[0], a[1], a[2], a[0]/a[2], a[1]/a[2]))
