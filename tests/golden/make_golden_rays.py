"""Generates tests/golden/rays.npz by running the REFERENCE'S OWN ray helpers on CPU (build container only):
get_rays_of_a_view / ndc_rays (lib/voxurf_fine.py:1001-1070) for several flag combinations and
get_training_rays_in_maskcache_sampling (:1127-1164) with the reference Voxurf model's hit_coarse_geo.
Same stubbing as make_golden.py; inputs come from voxurf_b200.synthetic.make_views (seeded), only outputs are stored.

    python tests/golden/make_golden_rays.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402
from voxurf_b200 import synthetic as S  # noqa: E402

T = MG.T


def main():
    MG._install_stubs()
    sys.path.insert(0, '/root/reference')
    import lib.voxurf_fine as vf
    out = MG.Snap()
    # ---- single views, flag combinations (S.RAY_CASES)
    for name, kw in S.RAY_CASES.items():
        H, W, K, c2w = S.make_view(seed=kw['seed'], H=kw['H'], W=kw['W'], inverse_y=kw['inverse_y'])
        ro, rd, vd = vf.get_rays_of_a_view(H, W, K, T(c2w), kw['ndc'], kw['inverse_y'], kw['flip_x'], kw['flip_y'],
                                           mode=kw['mode'])
        out.update({name + '_rays_o': ro.contiguous(), name + '_rays_d': rd, name + '_viewdirs': vd})
    # ---- training-ray gathering with the in-mask-cache filter
    G, C, Wd = 20, 6, 32
    sc = S.make_fine_scene(G, C, Wd, seed=3, mask_G=12)
    ckpt = '/tmp/voxurf_golden/mask_rays.tar'
    os.makedirs(os.path.dirname(ckpt), exist_ok=True)
    MG.write_mask_ckpt(sc, ckpt)
    m = MG.build_fine(vf, sc, ckpt)
    views = [S.make_view(seed=40 + i, H=h, W=w, inverse_y=False) for i, (h, w) in enumerate(S.TRAIN_VIEW_SIZES)]
    imgs = [T(S.make_image(h, w, seed=i)) for i, (h, w, _, _) in enumerate(views)]
    poses = torch.stack([T(v[3]) for v in views])
    HW = np.array([(v[0], v[1]) for v in views])
    Ks = np.stack([v[2] for v in views])
    rk = dict(near=0.3, far=6.0, bg=0.0, stepsize=0.5)
    rgb, ro, rd, vd, imsz = vf.get_training_rays_in_maskcache_sampling(imgs, poses, HW, Ks, False, False, False, False, m, rk)
    out.update(tr_rgb=rgb, tr_rays_o=ro, tr_rays_d=rd, tr_viewdirs=vd, tr_imsz=torch.tensor([int(n) for n in imsz]))
    np.savez_compressed(os.path.join(HERE, 'rays.npz'), **MG.npy(out))
    print('rays.npz', os.path.getsize(os.path.join(HERE, 'rays.npz')), {k: tuple(v.shape) for k, v in out.items() if k.startswith('tr_')})


if __name__ == '__main__':
    main()
