"""Attribute the warp-stall samples / executed instructions of an ncu SASS source page to CUDA source lines.
usage: ncu_lines.py <report.ncu-rep> <cubin> <mangled kernel name> [source.cu]   (cubin: cuobjdump -xelf all lib.so)"""
import csv, re, subprocess, sys
from collections import defaultdict
rep, cubin, kern = sys.argv[1:4]
srcfile = sys.argv[4] if len(sys.argv) > 4 else None
txt = subprocess.run(['nvdisasm', '-g', cubin], capture_output=True, text=True).stdout.splitlines()
seq, cur, on = [], None, False
for ln in txt:
    if ln.startswith('.text.') or re.match(r'\s*\.section\s+\.text\.', ln):
        on = kern in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File "(?:.*/)?([^/"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        seq.append((cur, m.group(2).strip()))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]
si, ins, ie = h.index('Warp Stall Sampling (All Samples)'), h.index('Source'), h.index('Instructions Executed')
data = []
for r in rows[hi + 1:]:
    try:
        data.append((int(r[si] or 0), r[ins], int(r[ie] or 0)))
    except Exception:
        pass
print('sass instrs', len(seq), 'ncu instrs', len(data))
by = defaultdict(lambda: [0, 0])
for (loc, sass), (smp, s_, ex) in zip(seq, data):
    by[loc][0] += smp; by[loc][1] += ex
tot = sum(v[0] for v in by.values()); totex = sum(v[1] for v in by.values())
src = open(srcfile).read().splitlines() if srcfile else []
for loc, (smp, ex) in sorted(by.items(), key=lambda x: -x[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 45]:
    if loc is None:
        continue
    f, l = loc
    text = src[l - 1].strip()[:100] if src and f == srcfile.split('/')[-1] and l <= len(src) else ''
    print(f'{100 * smp / tot:5.1f}% stall  {100 * ex / totex:5.1f}% exec  {f}:{l:4d}  {text}')
