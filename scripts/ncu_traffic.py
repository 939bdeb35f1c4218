"""Summarise an `ncu --set full` report for profiles/: per kernel (mean over the captured launches) duration, dram bytes
read / written, tensor-pipe activity, issue activity, L2 hit rate, registers.  Writes profiles/traffic.json keyed by the
hash of the CUDA sources (bench.py prints `roofline.traffic` only when the hash matches the build it runs).

    python scripts/ncu_traffic.py gpurun_out/prof.ncu-rep | gpurun_out/prof_raw.csv [profiles/name.md]
"""
import csv
import json
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv, argv = sys.argv[:1], sys.argv
import bench  # noqa: E402

rep = argv[1]
# (a report of ~80 launches is > 100 MB and cannot come back from the GPU box: scripts/profile_step.sh exports the raw page there)
out = open(rep).read() if rep.endswith('.csv') else subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
col = {c: i for i, c in enumerate(h)}
want = {'dur_us': 'gpu__time_duration.sum', 'dram_rd': 'dram__bytes_read.sum', 'dram_wr': 'dram__bytes_write.sum',
        'tensor_pct': 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'issue_pct': 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l2_hit_pct': 'lts__t_sector_hit_rate.pct', 'regs': 'launch__registers_per_thread', 'lts_bytes': 'lts__t_bytes.sum',
        'dram_pct': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1_pct': 'l1tex__throughput.avg.pct_of_peak_sustained_active'}
units = rows[1]


def num(r, key):
    c = want[key]
    if c not in col:
        return None
    v = r[col[c]].replace(',', '')
    try:
        x = float(v)
    except ValueError:
        return None
    u = units[col[c]].lower()
    if key in ('dram_rd', 'dram_wr', 'lts_bytes'):
        x *= {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
    if key == 'dur_us':
        x *= {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1, 'msecond': 1e3}.get(u, 1)
    return x


agg = defaultdict(list)
for r in rows[2:]:
    name = re.sub(r'\(.*', '', r[col['Kernel Name']])
    name = re.sub(r'<.*', '', name).replace('void ', '')
    agg[name].append({k: num(r, k) for k in want})
summary = {}
for name, ls in agg.items():
    summary[name] = {k: (sum(x[k] for x in ls if x[k] is not None) / max(1, sum(1 for x in ls if x[k] is not None))) for k in want}
    summary[name]['launches'] = len(ls)
traffic = {'src_sha': bench.source_sha(), 'report': os.path.basename(rep),
           'kernels': {k: v['dram_rd'] + v['dram_wr'] for k, v in summary.items() if v['dram_rd'] is not None}}
json.dump(traffic, open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w'), indent=1)
lines = ['| kernel | launches | time (us) | dram read / write (MB) | dram % | L2 hit % | L2 bytes (MB) | tensor pipe % | issue % | L1 % | regs |', '|---|---|---|---|---|---|---|---|---|---|---|']
f = lambda x, d=1: '-' if x is None else f'{x:.{d}f}'
for k, v in sorted(summary.items(), key=lambda kv: -(kv[1]['dur_us'] or 0)):
    lines.append(f"| `{k}` | {v['launches']} | {f(v['dur_us'])} | {f(v['dram_rd'] / 1e6 if v['dram_rd'] is not None else None)} / {f(v['dram_wr'] / 1e6 if v['dram_wr'] is not None else None)} | "
                 f"{f(v['dram_pct'])} | {f(v['l2_hit_pct'])} | {f(v['lts_bytes'] / 1e6 if v['lts_bytes'] is not None else None)} | {f(v['tensor_pct'])} | {f(v['issue_pct'])} | {f(v['l1_pct'])} | {f(v['regs'], 0)} |")
text = '\n'.join(lines)
print(text)
if len(argv) > 2:
    open(argv[2], 'a').write('\n' + text + '\n')
