"""Aggregate warp-stall samples of an .ncu-rep per CUDA source line:
   ncu -i X.ncu-rep --page source --csv --print-source sass,cuda > both.csv; python profiles/stall_by_line.py both.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, agg = None, None, {}
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 5 and r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    try:
        n = int(r[hdr.index('# Samples')])
        line = int(r[0])
    except ValueError:
        continue
    key = (cur, line, r[1].strip()[:110])
    ent = agg.setdefault(key, [0, {}])
    ent[0] += n
    for j, h in enumerate(hdr):
        if h.startswith('stall_') and 'Not Issued' not in h:
            try:
                ent[1][h[6:]] = ent[1].get(h[6:], 0) + int(r[j])
            except ValueError:
                pass
tot = sum(v[0] for v in agg.values())
print('total samples', tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(((c, s) for s, c in v[1].items()), reverse=True)[:3]
    print(f"{v[0]:5d} {100 * v[0] / max(tot, 1):5.1f}% {k[0]}:{k[1]:<4d} {k[2][:95]:95s} {st}")
