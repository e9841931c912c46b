"""Rank CUDA source lines by warp-stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
hdr = None
out = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        hdr = None
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[2] != "-":  # keep the per-source-line aggregate rows only (Address == "-")
        continue
    d = {}
    for k, v in zip(hdr, r):
        d.setdefault(k, v)
    try:
        s = int(d["# Samples"])
    except (KeyError, ValueError):
        continue
    out.append((s, cur, d["Line No"], r[1].strip()[:100], d))
tot = sum(o[0] for o in out)
print("total samples", tot)
out.sort(key=lambda x: -x[0])
keys = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_selected",
        "stall_branch_resolving", "stall_no_inst"]
for s, f, l, src, d in out[:top]:
    st = " ".join(f"{k[6:9]}={d.get(k, '')}" for k in keys if d.get(k, "0") not in ("0", ""))
    print(f"{100 * s / tot:5.1f}% {f}:{l} | {src} | {st}")
