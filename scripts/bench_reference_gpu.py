"""the REFERENCE'S GPU path timed on this B200 -- lib/voxurf_fine.py + lib/utils.Adam unmodified
(baseline/_ref/lib) on the reference's own CUDA kernels compiled for sm_100a (oracle/_ref) and stock ATen
(grid_sample / Linear / index_add), training-loop body of run.py:600-683, same scene / batch / step numbers as bench.py.
Also the same reference Python on this repository's shim (backend b200).  TEST / BASELINE infrastructure: uses oracle/.

    python scripts/bench_reference_gpu.py [--steps 30] [--warmup 10] [--backends ref,b200] > profiles/...json
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from oracle import ref_model as RM  # noqa: E402
from voxurf_b200 import synthetic as S  # noqa: E402
from voxurf_b200.trainer import FINE_TRAIN  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=30)
ap.add_argument('--warmup', type=int, default=10)
ap.add_argument('--backends', default='ref,b200')
ap.add_argument('--grid', type=int, default=256)
ap.add_argument('--k0-channels', type=int, default=12)
a = ap.parse_args()
G, C, N = a.grid, a.k0_channels, 8192
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
ckpt = '/tmp/vx_mask_bench.tar'
RM.write_mask_ckpt(ckpt, S.mask_density(100), float(np.log(1 / (1 - 1e-6) - 1)))
cfg_model = {k: v for k, v in S.FINE_CFG.items() if k != 'stepsize'}
rs = np.random.RandomState(0)
d1, d2 = S.fine_dims(C)
mlps = (S.mlp_init(rs, d1, 192, 4), S.mlp_init(rs, d2, 192, 4))
sdf = torch.from_numpy(S.sphere_sdf(G))
k0 = 0.1 * torch.randn(1, C, G, G, G, generator=torch.Generator().manual_seed(1234))
pool = [tuple(t.to(dev) for t in b) for b in B.ray_pool(8, N, 0)]
cfg = RM.Cfg(FINE_TRAIN)
res = {}
for backend in a.backends.split(','):
    ns = RM.load_lib(backend)
    m = RM.build_fine(ns, G, C, 192, ckpt, cfg_model, sdf, k0, mlps)
    opt = RM.make_optimizer(ns, m, cfg)
    gs = B.START_STEP
    for _ in range(a.warmup):
        RM.train_step(m, opt, cfg, B.RENDER_KW, pool[gs % 8], gs); gs += 1
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss, rr = RM.train_step(m, opt, cfg, B.RENDER_KW, pool[gs % 8], gs); gs += 1
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    res[backend] = {'ms_per_step': ms, 'rays_per_s': N / ms * 1e3, 'loss': float(loss), 'M4': int(rr['weights'].shape[0]),
                    'what': {'ref': "reference Python + reference CUDA kernels (oracle/_ref) + ATen + lib/utils.Adam",
                             'b200': "reference Python unmodified on voxurf_b200's shim modules (C ABI) + ATen + lib/utils.Adam"}[backend]}
    del m, opt
    torch.cuda.empty_cache()
print(json.dumps({'metric': B.METRIC, 'config': {'grid': G, 'k0_channels': C, 'rays': N, 'steps': a.steps, 'warmup': a.warmup,
                                                  'start_step': B.START_STEP}, 'baseline_B2': res}))
