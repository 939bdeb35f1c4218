"""Measure the issue rate of tcgen05.mma.kind::tf32 (M = 128, K = 8) for several widths / forms (development)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from voxurf_b200._lib import call
dev = 'cuda'
for M in (128, 64):
    for blocks in (1, 148):
        for form in (0, 1):
            row = []
            for N in (16, 64, 96, 128, 192, 208, 240):
                n_acc = 1
                cyc = torch.zeros(blocks, dtype=torch.int64, device=dev)
                call('vx_umma_rate', blocks, 64, M, N, form, n_acc, cyc)     # warm
                c1 = torch.zeros(blocks, dtype=torch.int64, device=dev)
                call('vx_umma_rate', blocks, 256, M, N, form, n_acc, c1)
                c2 = torch.zeros(blocks, dtype=torch.int64, device=dev)
                call('vx_umma_rate', blocks, 2304, M, N, form, n_acc, c2)
                torch.cuda.synchronize()
                per = (c2.double().mean() - c1.double().mean()) / 2048
                row.append('N=%d: %.0f clk (%.0f MAC/clk)' % (N, per, M * N * 8 / per))
            print('M', M, 'blocks', blocks, 'form', 'SS' if form == 0 else 'TS', '|', ' | '.join(row))

# where do the rows of an M = 64 accumulator land in TMEM?  one-hot A rows, B = ones, dump all 128 lanes
import numpy as np
def desc(lbo, sbo, layout=0):
    return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | 1 << 46 | layout << 61
N = 16
B_img = torch.zeros(8192, device=dev)
for n in range(N):
    for k in range(8):
        B_img[((k // 4) * 256 + (n // 8) * 128 + (n % 8) * 16 + (k % 4) * 4) // 4] = 1.0
idesc64 = (1 << 4) | (2 << 7) | (2 << 10) | ((N >> 3) << 17) | ((64 >> 4) << 24)
lanes = {}
for m in range(64):
    A_img = torch.zeros(8192, device=dev)
    A_img[((m // 8) * 128 + (m % 8) * 16) // 4] = 1.0      # K-major: element (m, k=0)
    D = torch.zeros(128, N, device=dev)
    call('vx_umma_probe', A_img, 8192, B_img, 8192, desc(2048, 128), desc(256, 128), idesc64, N, D)
    nz = torch.nonzero(D[:, 0]).flatten().tolist()
    lanes[m] = nz
print('M=64 accumulator: row -> TMEM lane(s):', {m: lanes[m] for m in (0, 1, 15, 16, 31, 32, 47, 48, 63)})
