"""Measure the issue rate of tcgen05.mma.kind::tf32 (M = 128, K = 8) for several widths / forms (development)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from voxurf_b200._lib import call
dev = 'cuda'
for blocks in (1, 148):
    for form in (0, 1):
        for n_acc in (1, 2):
            row = []
            for N in (16, 64, 96, 128, 192, 240):
                cyc = torch.zeros(blocks, dtype=torch.int64, device=dev)
                call('vx_umma_rate', blocks, 64, N, form, n_acc, cyc)     # warm
                c1 = torch.zeros(blocks, dtype=torch.int64, device=dev)
                call('vx_umma_rate', blocks, 256, N, form, n_acc, c1)
                c2 = torch.zeros(blocks, dtype=torch.int64, device=dev)
                call('vx_umma_rate', blocks, 2304, N, form, n_acc, c2)
                torch.cuda.synchronize()
                per = (c2.double().mean() - c1.double().mean()) / 2048
                row.append('N=%d: %.0f clk (%.0f MAC/clk)' % (N, per, 128 * N * 8 / per))
            print('blocks', blocks, 'form', 'SS' if form == 0 else 'TS', 'accumulators', n_acc, '|', ' | '.join(row))
