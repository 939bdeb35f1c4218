"""Print the kernels of the last full step of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
L = [(re.sub(r'\(.*', '', r[ki]), float(r[vi].replace(',', '')) / 1000) for r in rows[hdr + 1:] if len(r) > vi]
idx = [i for i, (n, _) in enumerate(L) if 'march_flags' in n]
for a, b in zip(idx[-4:-1], idx[-3:]):
    print('step: launches', b - a, 'sum us %.1f' % sum(t for _, t in L[a:b]))
if len(sys.argv) > 2:
    a, b = idx[-1 - int(sys.argv[2])], idx[-int(sys.argv[2])] if int(sys.argv[2]) > 0 else len(L)
else:
    a, b = idx[-2], idx[-1]
for n, t in L[a:b]:
    print(f'{t:9.1f}  {n[:70]}')
