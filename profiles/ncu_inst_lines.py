"""Rank CUDA source lines by executed warp instructions (ncu --page source --csv --print-source cuda,sass)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0  # e.g. warps * steps: prints instructions per unit
cur = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "Function Name" or r[2] != "-": continue
    d = {}
    for k, v in zip(hdr, r): d.setdefault(k, v)
    try: n = int(d.get("Instructions Executed", "0"))
    except ValueError: continue
    out.append((n, cur, d["Line No"], r[1].strip()[:110]))
tot = sum(o[0] for o in out)
print("total warp instructions", tot, "per unit", tot / units)
out.sort(key=lambda x: -x[0])
for n, f, l, src in out[:top]:
    print(f"{100 * n / tot:5.1f}% {n / units:8.1f} {f}:{l} | {src}")
