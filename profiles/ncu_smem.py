"""Shared-memory wavefronts per source line / instruction type from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur = None; hdr = None; line = None; src = {}; agg = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; hdr = None; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None: continue
    if r[2] == '-':
        line = r[0]; src[(cur, line)] = r[1].strip(); continue
    d = {}
    for k, v in zip(hdr, r): d.setdefault(k, v)
    if r[0] not in ('', '-'): line = r[0]
    toks = r[3].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
    if not (op.startswith('LDS') or op.startswith('STS') or op.startswith('LDGSTS')): continue
    try: n = int(d['Instructions Executed']); wf = int(d['L1 Wavefronts Shared']); ideal = int(d['L1 Wavefronts Shared Ideal'])
    except Exception: continue
    a = agg.setdefault((cur, line, op.split('.')[0] + '.' + ('128' if '128' in op else '64' if '64' in op else '32')), [0, 0, 0])
    a[0] += n; a[1] += wf; a[2] += ideal
tot = sum(a[1] for a in agg.values()); print('total wf', tot)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*a[1]/tot:5.1f}% {k[0]}:{k[1]:>4} {k[2]:10s} instr={a[0]:>10} wf={a[1]:>11} wf/instr={a[1]/max(a[0],1):.2f} ideal={a[2]/max(a[0],1):.2f} | {src.get((k[0],k[1]),'')[:90]}")
