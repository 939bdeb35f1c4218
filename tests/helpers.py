"""Shared test helpers: build oracle / product models from the seeded synthetic scenes."""
import os

import numpy as np
import torch

from oracle import voxurf_ref as R
from voxurf_b200 import synthetic as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def oracle_mask_cache(sc, thres=1e-3):
    return dict(density=R.mask_cache_density(T(sc['mask_density'])), xyz_min=torch.tensor([-1., -1., -1.]),
                xyz_max=torch.tensor([1., 1., 1.]), act_shift=sc['mask_act_shift'],
                voxel_size_ratio=sc['mask_voxel_size_ratio'], thres=thres)


def _layers(ls, requires_grad):
    return [(T(W).clone().requires_grad_(requires_grad), T(b).clone().requires_grad_(requires_grad)) for W, b in ls]


def oracle_fine_model(sc, requires_grad=True, apply_nonempty=True):
    G = sc['G']
    xyz_min, xyz_max = torch.tensor([-1., -1., -1.]), torch.tensor([1., 1., 1.])
    voxel_size = ((xyz_max - xyz_min).prod() / (G ** 3)).pow(1 / 3)
    cfg = S.FINE_CFG
    m = dict(xyz_min=xyz_min, xyz_max=xyz_max, voxel_size=voxel_size,
             posfreq=torch.FloatTensor([2 ** i for i in range(cfg['posbase_pe'])]),
             viewfreq=torch.FloatTensor([2 ** i for i in range(cfg['viewbase_pe'])]),
             k_posfreq=torch.FloatTensor([2 ** i for i in range(cfg['k_posbase_pe'])]),
             k_viewfreq=torch.FloatTensor([2 ** i for i in range(cfg['k_viewbase_pe'])]),
             grad_feat=cfg['grad_feat'], use_grad_norm=cfg['use_grad_norm'], center_sdf=cfg['center_sdf'],
             k_center_sdf=cfg['k_center_sdf'], k_res=cfg['k_res'], fast_color_thres=cfg['fast_color_thres'],
             s_ratio=cfg['s_ratio'], s_start=cfg['s_start'], step_start=0, s_val=cfg['s_start'], smooth_kernel=None)
    sdf = T(sc['sdf']).clone()
    if 'mask_density' in sc:
        m['mask_cache'] = oracle_mask_cache(sc, cfg['mask_cache_thres'])
        m['nonempty_mask'] = R.nonempty_mask(m['mask_cache'], xyz_min, xyz_max, (G, G, G))
        if apply_nonempty:
            sdf[~m['nonempty_mask']] = 1   # voxurf_fine.py:367
    else:
        m['mask_cache'] = None
    m['sdf'] = sdf.requires_grad_(requires_grad)
    m['k0'] = T(sc['k0']).clone().requires_grad_(requires_grad)
    m['rgbnet'] = _layers(sc['rgbnet'], requires_grad)
    m['k_rgbnet'] = _layers(sc['k_rgbnet'], requires_grad)
    return m


def oracle_coarse_model(sc, requires_grad=True, apply_nonempty=True):
    G = sc['G']
    xyz_min, xyz_max = torch.tensor([-1., -1., -1.]), torch.tensor([1., 1., 1.])
    voxel_size = ((xyz_max - xyz_min).prod() / (G ** 3)).pow(1 / 3)
    cfg = S.COARSE_CFG
    m = dict(xyz_min=xyz_min, xyz_max=xyz_max, voxel_size=voxel_size,
             posfreq=torch.FloatTensor([2 ** i for i in range(cfg['posbase_pe'])]),
             viewfreq=torch.FloatTensor([2 ** i for i in range(cfg['viewbase_pe'])]),
             fast_color_thres=cfg['fast_color_thres'], s_ratio=cfg['s_ratio'], s_start=cfg['s_start'], step_start=0,
             s_val=cfg['s_start'], smooth_kernel=R.gaussian_kernel3d(cfg['smooth_ksize'], cfg['smooth_sigma']))
    sdf = T(sc['sdf']).clone()
    if 'mask_density' in sc:
        m['mask_cache'] = oracle_mask_cache(sc, cfg['mask_cache_thres'])
        m['nonempty_mask'] = R.nonempty_mask(m['mask_cache'], xyz_min, xyz_max, (G, G, G))
        if apply_nonempty:
            sdf[~m['nonempty_mask']] = 1
    else:
        m['mask_cache'] = None
    m['sdf'] = sdf.requires_grad_(requires_grad)
    m['k0'] = T(sc['k0']).clone().requires_grad_(requires_grad)
    m['rgbnet'] = _layers(sc['rgbnet'], requires_grad)
    return m


# ------------------------------------------------------------------------------------------------
# product-side models from the same scenes
# ------------------------------------------------------------------------------------------------
def mask_cache_state(sc):
    return {'MaskCache_kwargs': {'xyz_min': [-1., -1., -1.], 'xyz_max': [1., 1., 1.], 'act_shift': sc['mask_act_shift'],
                                 'voxel_size_ratio': sc['mask_voxel_size_ratio'], 'nearest': False},
            'model_state_dict': {'density': T(sc['mask_density'])}}


def set_mlp(seq, layers):
    lin = [m for m in seq.modules() if isinstance(m, torch.nn.Linear)]
    assert len(lin) == len(layers)
    for m, (W, b) in zip(lin, layers):
        m.weight.data = T(W).clone()
        m.bias.data = T(b).clone()


def product_fine_model(sc, device='cuda', apply_nonempty=True, k0_channels_last=False):
    from voxurf_b200 import voxurf_fine as VF
    G = sc['G']
    cfg = {k: v for k, v in S.FINE_CFG.items() if k != 'stepsize'}
    m = VF.Voxurf(xyz_min=[-1., -1., -1.], xyz_max=[1., 1., 1.], num_voxels=G ** 3, num_voxels_base=G ** 3,
                  rgbnet_dim=sc['C'], rgbnet_width=sc['width'], k0_channels_last=k0_channels_last,
                  mask_cache_state=mask_cache_state(sc) if 'mask_density' in sc else None, **cfg)
    assert tuple(int(w) for w in m.world_size) == (G, G, G)
    m.sdf.grid.data = T(sc['sdf']).clone()
    k0 = T(sc['k0']).clone()
    m.k0.grid.data = k0.contiguous(memory_format=torch.channels_last_3d) if k0_channels_last else k0
    set_mlp(m.rgbnet, sc['rgbnet'])
    set_mlp(m.k_rgbnet, sc['k_rgbnet'])
    m = m.to(device)
    if m.mask_cache is not None and apply_nonempty:
        m._set_nonempty_mask()
    return m


def product_coarse_model(sc, device='cuda', apply_nonempty=True, k0_channels_last=False):
    from voxurf_b200 import voxurf_coarse as VC
    G = sc['G']
    cfg = {k: v for k, v in S.COARSE_CFG.items() if k != 'stepsize'}
    cfg['rgbnet_dim'], cfg['rgbnet_width'] = sc['C'], sc['width']
    m = VC.Voxurf(xyz_min=[-1., -1., -1.], xyz_max=[1., 1., 1.], num_voxels=G ** 3, num_voxels_base=G ** 3,
                  rgbnet_direct=True, k0_channels_last=k0_channels_last,
                  mask_cache_state=mask_cache_state(sc) if 'mask_density' in sc else None, **cfg)
    assert tuple(int(w) for w in m.world_size) == (G, G, G)
    m.sdf.grid.data = T(sc['sdf']).clone()
    k0 = T(sc['k0']).clone()
    m.k0.grid.data = k0.contiguous(memory_format=torch.channels_last_3d) if k0_channels_last else k0
    set_mlp(m.rgbnet, sc['rgbnet'])
    m = m.to(device)
    if m.mask_cache is not None and apply_nonempty:
        m._set_nonempty_mask()
    return m
