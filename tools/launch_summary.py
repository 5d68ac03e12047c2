"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X): python tools/launch_summary.py X [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hd = rows[h]
kn, mv, un = hd.index("Kernel Name"), hd.index("Metric Value"), hd.index("Metric Unit")
t, n = collections.Counter(), collections.Counter()
for r in rows[h + 1:]:
    if len(r) > mv:
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1., "s": 1e3}.get(r[un], 1e-6)
        k = r[kn][:70]
        t[k] += v
        n[k] += 1
tot = sum(t.values())
print("total %.3f ms in %d launches" % (tot, sum(n.values())))
for k, v in t.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 20):
    print("%-70s %9.3f ms %6d launches %5.1f%%  %.4f ms each" % (k, v, n[k], 100 * v / tot, v / n[k]))
