"""Micro-benchmark of the tcgen05 MLP launches at the fine stage's shape (both networks, ~45 k rows):
forward pair, dX pair, batched dW -- CUDA events, medians.   python scripts/bench_mlp.py [rows] [iters]"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxurf_b200.mlp import FlatMLP, prepare_chains, run_chain_jobs, run_dw_batch  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 45137
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = 'cuda'
torch.manual_seed(0)


def net(d):
    return nn.Sequential(nn.Linear(d, 192), nn.ReLU(True), nn.Sequential(nn.Linear(192, 192), nn.ReLU(True)),
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


def fwd():
    run_chain_jobs([m1.forward_job(X1, o1, True), m2.forward_job(X2, o2, True, patch=(o1, 57, 3, 0))], n, cap, m1.done)


def bwd():
    run_chain_jobs([m2.backward_job(d2, dX2), m1.backward_job(d1, dX1)], n, cap, None)


def dw():
    run_dw_batch([m2, m1])


def timeit(f, reps=8):
    """`reps` launches captured in one CUDA graph and replayed: pure device time, no host launch overhead"""
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            f()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / reps)
    return float(np.median(ts)), float(np.min(ts))


fl = (m1.tc_fwd.flops_per_row() + m2.tc_fwd.flops_per_row()) * rows
for name, f, flops in (('forward pair', fwd, fl), ('dX pair', bwd, (m1.tc_bwd.flops_per_row() + m2.tc_bwd.flops_per_row()) * rows), ('dW batch', dw, fl)):
    med, mn = timeit(f)
    print(f'{name:14s} median {med:7.1f} us  min {mn:7.1f} us   {flops / med / 1e6:6.1f} TFLOP/s fp32-equivalent ({3 * flops / med / 1e6:6.1f} issued TF32)')
