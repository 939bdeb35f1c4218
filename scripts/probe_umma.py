"""Reverse-engineer how tcgen05.mma (kind::tf32) reads an MN-major shared-memory operand.

For every float position p of the A image a one-hot image is multiplied with a K-major B whose column k holds k + 1:
D[m][n] = k + 1 at the (m, k) that position p feeds.  Writes gpurun_out/umma_map.json: {variant: [[m, k] or None per p]}.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from voxurf_b200._lib import call

dev = 'cuda'
N = 16


def desc(lbo, sbo, layout):
    return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | 1 << 46 | layout << 61


def idesc(M, Nn, a_mn, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | (a_mn << 15) | (b_mn << 16) | ((Nn >> 3) << 17) | ((M >> 4) << 24)


def kmajor_image(V, lbo, sbo):   # V (MN, 8) -> floats
    MN = V.shape[0]
    img = torch.zeros(8192)
    for mn in range(MN):
        for k in range(8):
            img[((k // 4) * lbo + (mn // 8) * sbo + (mn % 8) * 16 + (k % 4) * 4) // 4] = V[mn, k]
    return img


B = torch.zeros(N, 8)
for k in range(8):
    B[:, k] = k + 1
B_img = kmajor_image(B, 256, 128).to(dev)
db = desc(256, 128, 0)

# sanity: K-major A
A = torch.randint(-4, 5, (128, 8)).float()
D = torch.zeros(128, N, device=dev)
call('vx_umma_probe', kmajor_image(A, 2048, 128).to(dev), 8192, B_img, 8192, desc(2048, 128, 0), db, idesc(128, N, 0, 0), N, D)
print('K-major sanity err', (D.cpu() - A @ B.t()).abs().max().item())

out = {}
variants = [('L1_lbo512_sbo2048', 1, 512, 2048), ('L1_lbo2048_sbo512', 1, 2048, 512), ('L0_lbo4096_sbo128', 0, 4096, 128),
            ('L0_lbo128_sbo4096', 0, 128, 4096), ('L2_lbo1024_sbo4096', 2, 1024, 4096), ('L2_lbo4096_sbo1024', 2, 4096, 1024),
            ('L4_lbo512_sbo2048', 4, 512, 2048), ('L6_lbo512_sbo2048', 6, 512, 2048)]
P = 2048
for name, layout, lbo, sbo in variants:
    da = desc(lbo, sbo, layout)
    idc = idesc(128, N, 1, 0)
    res = []
    eye = torch.zeros(8192, device=dev)
    for p in range(P):
        eye.zero_()
        eye[p] = 1.0
        D.zero_()
        call('vx_umma_probe', eye, 8192, B_img, 8192, da, db, idc, N, D)
        col = D[:, 0]
        nz = torch.nonzero(col).flatten().tolist()
        if len(nz) == 0:
            res.append(None)
        else:
            res.append([[m, int(round(float(col[m]))) - 1] for m in nz])
    hit = sum(r is not None for r in res)
    print(name, 'positions feeding the MMA:', hit, 'first 12:', res[:12])
    out[name] = res
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/umma_map.json', 'w'))
