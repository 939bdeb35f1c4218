"""Pipeline trace of CTA 0 of the chain kernel (needs the -DMC_TRACE build:
VX_NVCC_FLAGS=-DMC_TRACE VX_OBJ_DIR=_obj_trace VX_SO=$PWD/voxurf_b200/libvoxurf_b200_trace.so python -m voxurf_b200.build;
run with VX_SO set to that library).  Prints the event log of one forward-pair and one dX-pair launch in cycles."""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxurf_b200._lib import call  # noqa: E402
from voxurf_b200.mlp import FlatMLP, prepare_chains, run_chain_jobs  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 45137
dev = 'cuda'
torch.manual_seed(0)
net = lambda d: nn.Sequential(nn.Linear(d, 192), nn.ReLU(True), nn.Sequential(nn.Linear(192, 192), nn.ReLU(True)),
                              nn.Sequential(nn.Linear(192, 192), nn.ReLU(True)), nn.Linear(192, 3)).to(dev)
cap = (int(rows * 1.35) + 4095) // 4096 * 4096
m1, m2 = FlatMLP(net(79), 80, 79), FlatMLP(net(60), 64, 60)
m1.alloc(cap), m2.alloc(cap)
X1 = torch.randn(cap, 80, device=dev); X2 = torch.randn(cap, 64, device=dev)
o1, o2 = torch.zeros(cap, 3, device=dev), torch.zeros(cap, 3, device=dev)
d1, d2 = torch.randn(cap, 3, device=dev) * 1e-3, torch.randn(cap, 3, device=dev) * 1e-3
dX1, dX2 = torch.zeros_like(X1), torch.zeros_like(X2)
n = torch.tensor([rows], dtype=torch.int32, device=dev)
m1._n = m2._n = n
prepare_chains(m1.chains() + m2.chains())
fwd = lambda: run_chain_jobs([m1.forward_job(X1, o1, True), m2.forward_job(X2, o2, True, patch=(o1, 57, 3, 0))], n, cap, m1.done)
bwd = lambda: run_chain_jobs([m2.backward_job(d2, dX2), m1.backward_job(d1, dX1)], n, cap, None)
NAMES = {1: 'acc ready', 2: 'acc wait begins', 10: 'chunk[0] done', 11: 'chunk[1] done', 12: 'chunk[2] done', 20: 'staged next',
         30: 'item done', 40: 'MMA: chunk[0] acquired', 41: 'MMA: chunk[1] acquired', 42: 'MMA: chunk[2] acquired', 50: 'MMA: layer committed', 70: '  visit begins', 71: '  acc columns in regs', 72: '  math done', 73: '  A/img stores issued', 74: '  tmem st done', 75: '  fences done', 60: 'MMA: weights wait', 61: 'MMA: weights in'}
for name, f in (('forward pair', fwd), ('dX pair', bwd)):
    for _ in range(3):
        f()
    buf = torch.zeros(4096, dtype=torch.int64, device=dev)
    call('vx_mlp_trace_set', buf)
    f()
    torch.cuda.synchronize()
    call('vx_mlp_trace_set', None)
    b = buf.cpu().tolist()
    ev = sorted([(v >> 8, v & 255, 'row') for v in b[:2048] if v] + [(v >> 8, v & 255, 'mma') for v in b[2048:] if v])
    t0 = ev[0][0]
    print(f'==== {name}: {len(ev)} events, {ev[-1][0] - t0} cycles')
    last = {'row': t0, 'mma': t0}
    for t, c, who in ev[:int(os.environ.get('TRACE_N', 90))]:
        print(f'{t - t0:8d}  (+{t - last[who]:6d})  {"    " if who == "row" else "                                "}{NAMES.get(c, c)}')
        last[who] = t
