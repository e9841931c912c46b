"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`) per kernel."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[h]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e6 if r[ui] == "ns" else (v / 1e3 if r[ui] == "us" else v)
    a = agg.setdefault(r[ki], [0.0, 0])
    a[0] += v; a[1] += 1
tot = sum(a[0] for a in agg.values())
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{100 * t / tot:6.2f}%  {t:10.3f} ms total  x{n:<4d} {t / n:9.3f} ms each  {k[:90]}")
