"""Top stalled SASS instructions of an `ncu --page source --csv` export (one kernel)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
si = h.index('# Samples'); src = h.index('Source'); ex = h.index('Instructions Executed')
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
body = [r for r in rows[2:] if len(r) > si]
tot = sum(int(r[si]) for r in body)
print('total samples', tot, 'instructions', len(body))
agg = {}
for i, c in stall_cols:
    agg[c] = sum(int(r[i]) for r in body)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for k, r in sorted(enumerate(body), key=lambda kr: -int(kr[1][si]))[:n]:
    st = sorted(((int(r[i]), c) for i, c in stall_cols), reverse=True)[:2]
    print(f'{k:5d} {int(r[si]):6d} {100*int(r[si])/tot:5.1f}%  ex={r[ex]:>7}  {r[src].strip()[:60]:60s} {st}')
