"""Coarse-stage training step (SURVEY 8d config 2 shape: 96^3 SDF + 12-ch k0, rgbnet 57->128^2->3, per-iteration 5^3
smoothing, 8192 rays) through the drop-in autograd path (Voxurf coarse mirror + Trainer).  Prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from voxurf_b200 import synthetic as S
from voxurf_b200.trainer import COARSE_TRAIN, Trainer
from tests.helpers import T, product_coarse_model

dev = 'cuda'
G, n_rays, steps, warm = 96, 8192, 30, 6
sc = S.make_coarse_scene(G, 12, 128, seed=0, mask_G=48)
m = product_coarse_model(sc)   # channel-major k0: the coarse stage regularises k0 with the autograd-form TV
tr = Trainer(m, COARSE_TRAIN, dict(near=0.3, far=6.0, bg=0.0, stepsize=0.5), zero_grad_in_step=False)
pool = []
for b in range(4):
    o, d, v = S.make_rays(n_rays, seed=900 + b)
    pool.append(tuple(T(x).to(dev) for x in (o, d, v, S.make_target(v, seed=b))))
for i in range(warm):
    tr.step(*pool[i % 4], global_step=1 + i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    loss, ret = tr.step(*pool[i % 4], global_step=1 + warm + i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({'workload': 'voxurf_coarse fwd+bwd+TV+Adam, 96^3, 12-ch k0, 8192 rays, drop-in autograd path', 'ms_per_step': ms,
                  'iters_per_s': 1000 / ms, 'rays_per_s': n_rays * 1000 / ms, 'M4': int(ret['weights'].shape[0])}))
