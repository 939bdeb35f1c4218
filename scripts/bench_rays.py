"""Time the caller-side ray gathering (get_training_rays_in_maskcache_sampling) on a DTU-shaped set of views:
49 views of 800x600, fine model with a 100^3 mask cache.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from voxurf_b200 import synthetic as S
from voxurf_b200 import voxurf_fine as vf
from tests.helpers import T, product_fine_model

dev = 'cuda'
n_views, H, W = int(os.environ.get('VIEWS', 49)), 600, 800
sc = S.make_fine_scene(64, 6, 32, seed=3, mask_G=100)
m = product_fine_model(sc)
views = [S.make_view(seed=100 + i, H=H, W=W) for i in range(n_views)]
imgs = [torch.rand(H, W, 3, device=dev) for _ in range(n_views)]
poses = torch.stack([T(v[3]) for v in views]).to(dev)
HW = np.array([(H, W)] * n_views); Ks = np.stack([v[2] for v in views])
rk = dict(near=0.3, far=6.0, bg=0.0, stepsize=0.5)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    out = vf.get_training_rays_in_maskcache_sampling(imgs, poses, HW, Ks, False, False, False, False, m, rk)
    torch.cuda.synchronize(); t1 = time.time()
print(json.dumps({'workload': 'get_training_rays_in_maskcache_sampling, %d views %dx%d' % (n_views, W, H), 'seconds': t1 - t0,
                  'rays_in': n_views * H * W, 'rays_kept': int(out[0].shape[0]), 'M_rays_per_s': n_views * H * W / (t1 - t0) / 1e6}))
