"""Shared-memory wavefronts per source line from an ncu report: python tools/ncu_smem_by_line.py report.ncu-rep [launch_index]
(ncu --page source --csv; columns 'L1 Wavefronts Shared', '... Ideal', '... Excessive')."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
blocks = out.split('"File Path"')
blk = '"File Path"' + blocks[1 + which]
lines = blk.splitlines()
hdr_i = next(i for i, l in enumerate(lines) if l.startswith('"Line No"'))
rd = csv.reader(io.StringIO("\n".join(lines[hdr_i:])))
hdr = next(rd)
col = {n: i for i, n in enumerate(hdr)}
iw, ii, ie, ix = col["L1 Wavefronts Shared"], col["L1 Wavefronts Shared Ideal"], col["L1 Wavefronts Shared Excessive"], col["Instructions Executed"]
agg = collections.defaultdict(lambda: [0, 0, 0, 0, ""])
cur_line, cur_src = None, ""
for row in rd:
    if len(row) < len(hdr): continue
    if row[0] != "":
        cur_line, cur_src = row[0], row[1]
    def f(v):
        try: return float(v)
        except ValueError: return 0.
    a = agg[cur_line]
    a[0] += f(row[iw]); a[1] += f(row[ii]); a[2] += f(row[ie]); a[3] += f(row[ix]) if row[2] != "" else 0; a[4] = cur_src
tot = sum(a[0] for a in agg.values())
print("total wavefronts %.3g  ideal %.3g  excessive %.3g" % (tot, sum(a[1] for a in agg.values()), sum(a[2] for a in agg.values())))
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    if a[0] == 0: break
    print("%6s  wavefronts %5.1f%%  ideal %.3g  excessive %.3g  inst %.3g | %s" % (ln, 100 * a[0] / tot, a[1], a[2], a[3], a[4].strip()[:90]))
