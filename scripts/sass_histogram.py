"""Opcode histogram of the Blackwell-specific instructions in libvoxurf_b200.so, per kernel: UTCHMMA / UTCQMMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit), SYNCS (mbarrier), RED / ATOM.
    python scripts/sass_histogram.py > profiles/r02_sass_histogram.md"""
import os
import re
import subprocess
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'voxurf_b200', 'libvoxurf_b200.so')
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
KEYS = ['UTCHMMA', 'UTCQMMA', 'UTCMMA', 'LDTM', 'STTM', 'UBLKCP', 'UTCBAR', 'UTCATOMSWS', 'SYNCS', 'RED', 'ATOM', 'STG', 'LDG', 'STS', 'LDS', 'SHFL', 'MATCH', 'VOTE']
per = defaultdict(Counter)
cur = None
for ln in txt.splitlines():
    m = re.search(r'Function : (\S+)', ln)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', ln)
    if m and cur:
        op = m.group(1).split('.')[0]
        per[cur]['total'] += 1
        for k in KEYS:
            if op.startswith(k):
                per[cur][k] += 1
                break
demangle = lambda s: subprocess.run(['c++filt', s], capture_output=True, text=True).stdout.strip().split('(')[0]
print('| kernel | SASS instrs | ' + ' | '.join(KEYS) + ' |')
print('|---|---|' + '---|' * len(KEYS))
for fn, c in sorted(per.items(), key=lambda kv: -kv[1]['total']):
    print(f'| `{demangle(fn)[:60]}` | {c["total"]} | ' + ' | '.join(str(c[k]) if c[k] else '' for k in KEYS) + ' |')
