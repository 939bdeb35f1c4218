"""Seeded synthetic scenes shared by the golden-vector generator, the tests and bench.py.

Everything here is numpy `RandomState` driven (stable across numpy versions) so that the tests can
regenerate the exact inputs the committed fixtures were produced from without storing them.
Shapes follow SURVEY.md section 8(d): analytic-sphere SDF on the reference's `np.mgrid` lattice
(lib/voxurf_fine.py:85-86), cameras on a sphere of radius 3 looking into the object, unnormalised
`rays_d`, `viewdirs = rays_d/|rays_d|`.
"""
import math

import numpy as np


def lattice_radius(G):
    x, y, z = np.mgrid[-1.0:1.0:G * 1j, -1.0:1.0:G * 1j, -1.0:1.0:G * 1j]
    return (x ** 2 + y ** 2 + z ** 2) ** 0.5


def sphere_sdf(G, radius=0.5, reduce=0.3, noise=0.0, seed=0):
    """(1,1,G,G,G) float32 SDF of a sphere, divided by `sdf_reduce` like the fine stage (fine.py:29)."""
    sdf = (lattice_radius(G) - radius) / reduce
    if noise > 0:
        sdf = sdf + noise * np.random.RandomState(seed).standard_normal(sdf.shape)
    return sdf.astype(np.float32)[None, None]


def mask_density(G, radius=0.62, sharp=200.0):
    """(1,1,G,G,G) raw coarse density whose activated alpha crosses 1e-3 near `radius`."""
    return ((radius - lattice_radius(G)) * sharp).astype(np.float32)[None, None]


def make_rays(n, seed=777, r_cam=3.0, r_target=0.5, jitter_len=True):
    """-> rays_o, rays_d, viewdirs (n,3) float32.  Origins uniform on the sphere r_cam, each ray aimed at
    a point uniform in the ball r_target (every ray hits the object, as the in_maskcache sampler
    guarantees, lib/voxurf_fine.py:1127-1164); |rays_d| != 1 like pixel-space directions."""
    rs = np.random.RandomState(seed)
    o = rs.standard_normal((n, 3))
    o = o / np.linalg.norm(o, axis=1, keepdims=True) * r_cam
    t = rs.standard_normal((n, 3))
    t = t / np.linalg.norm(t, axis=1, keepdims=True) * r_target * rs.uniform(0, 1, (n, 1)) ** (1 / 3)
    d = t - o
    v = d / np.linalg.norm(d, axis=1, keepdims=True)
    if jitter_len:
        d = v * rs.uniform(0.8, 1.3, (n, 1))
    return o.astype(np.float32), d.astype(np.float32), v.astype(np.float32)


def make_target(viewdirs, seed=1):
    rs = np.random.RandomState(seed)
    return (0.5 + 0.5 * np.sin(3 * viewdirs + rs.uniform(0, 6, (1, 3)))).astype(np.float32)


def linear_init(rs, fan_out, fan_in, zero_bias=False):
    b = 1.0 / math.sqrt(fan_in)
    W = rs.uniform(-b, b, (fan_out, fan_in)).astype(np.float32)
    bias = np.zeros(fan_out, np.float32) if zero_bias else rs.uniform(-b, b, (fan_out,)).astype(np.float32)
    return W, bias


def mlp_init(rs, dim0, width, depth, out=3):
    """Linear(dim0,width), (depth-2) x Linear(width,width), Linear(width,out) with zero last bias
    (lib/voxurf_fine.py:132-149)."""
    dims = [dim0] + [width] * (depth - 1) + [out]
    return [linear_init(rs, dims[i + 1], dims[i], zero_bias=(i == len(dims) - 2)) for i in range(len(dims) - 1)]


FINE_CFG = dict(posbase_pe=5, viewbase_pe=1, k_posbase_pe=5, k_viewbase_pe=1, k_res=True,
                rgbnet_depth=4, k_rgbnet_depth=4, k_grad_feat=(1.0,), k_sdf_feat=(), center_sdf=True,
                k_center_sdf=False, grad_feat=(0.5, 1.0, 1.5, 2.0), sdf_feat=(0.5, 1.0, 1.5, 2.0),
                use_grad_norm=True, stepsize=0.5, s_ratio=50, s_start=0.05, fast_color_thres=1e-4,
                mask_cache_thres=1e-3, alpha_init=1e-2)   # configs/dtu_e2e/fine.py:61-88

COARSE_CFG = dict(posbase_pe=5, viewbase_pe=1, rgbnet_depth=3, geo_rgb_dim=3, smooth_ksize=5,
                  smooth_sigma=0.8, s_ratio=50, s_start=0.2, stepsize=0.5, fast_color_thres=1e-4,
                  mask_cache_thres=1e-3, alpha_init=1e-2, rgbnet_dim=12, rgbnet_width=128)  # dtu_e2e/coarse.py:61-76


def fine_dims(C, cfg=FINE_CFG):
    L = len(cfg['grad_feat'])
    dim0 = (3 + 3 * cfg['posbase_pe'] * 2) + (3 + 3 * cfg['viewbase_pe'] * 2) + 3 * L + 6 * L + (1 if cfg['center_sdf'] else 0)
    k_dim0 = (3 + 3 * cfg['k_posbase_pe'] * 2) + (3 + 3 * cfg['k_viewbase_pe'] * 2) + C + 3 + 3 * len(cfg['k_grad_feat'])
    return dim0, k_dim0


def make_fine_scene(G, C=6, width=192, seed=0, mask_G=None, sdf_noise=0.01, with_mask=True):
    """Parameters of a fine-stage model as numpy arrays."""
    rs = np.random.RandomState(seed)
    dim0, k_dim0 = fine_dims(C)
    sc = dict(G=G, C=C, width=width,
              sdf=sphere_sdf(G, noise=sdf_noise, seed=seed + 1),
              k0=(0.1 * rs.standard_normal((1, C, G, G, G))).astype(np.float32),
              rgbnet=mlp_init(rs, dim0, width, FINE_CFG['rgbnet_depth']),
              k_rgbnet=mlp_init(rs, k_dim0, width, FINE_CFG['k_rgbnet_depth']))
    if with_mask:
        mg = mask_G or max(8, (G * 100) // 256)
        sc['mask_density'] = mask_density(mg)
        sc['mask_act_shift'] = float(np.log(1 / (1 - 1e-6) - 1))   # DVGO coarse alpha_init=1e-6
        sc['mask_voxel_size_ratio'] = 1.0
    return sc


def make_coarse_scene(G, C=12, width=128, seed=0, mask_G=None, sdf_noise=0.01, with_mask=True):
    rs = np.random.RandomState(seed)
    dim0 = (3 + 3 * 5 * 2) + (3 + 3 * 1 * 2) + C + 3
    sc = dict(G=G, C=C, width=width,
              sdf=sphere_sdf(G, reduce=1.0, noise=sdf_noise, seed=seed + 1),
              k0=(0.1 * rs.standard_normal((1, C, G, G, G))).astype(np.float32),
              rgbnet=mlp_init(rs, dim0, width, COARSE_CFG['rgbnet_depth']))
    if with_mask:
        mg = mask_G or max(8, G)
        sc['mask_density'] = mask_density(mg)
        sc['mask_act_shift'] = float(np.log(1 / (1 - 1e-6) - 1))
        sc['mask_voxel_size_ratio'] = 1.0
    return sc


# ------------------------------------------------------------------------------------------------
# synthetic views (pinhole cameras on a sphere looking at the origin) for the ray-generation path
# ------------------------------------------------------------------------------------------------
def make_view(seed, H, W, inverse_y=False, r_cam=3.0, fov_scale=1.1):
    """-> H, W, K (3,3) float32, c2w (4,4) float32.  OpenGL convention (camera looks along -z, y up) unless inverse_y
    (OpenCV: looks along +z, y down), matching the two branches of lib/voxurf_fine.py:1020-1023."""
    rs = np.random.RandomState(seed)
    p = rs.standard_normal(3)
    p = p / np.linalg.norm(p) * r_cam
    fwd = -p / np.linalg.norm(p)
    up0 = np.array([0.0, 0.0, 1.0]) if abs(fwd[2]) < 0.9 else np.array([0.0, 1.0, 0.0])
    right = np.cross(fwd, up0)
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    R = np.stack([right, -up, fwd], 1) if inverse_y else np.stack([right, up, -fwd], 1)
    c2w = np.eye(4, dtype=np.float32)
    c2w[:3, :3] = R.astype(np.float32)
    c2w[:3, 3] = p.astype(np.float32)
    focal = fov_scale * r_cam * max(H, W) / 2.0          # the unit-ish object fills most of the frame
    K = np.array([[focal, 0, W / 2.0 - 0.3], [0, focal * 1.01, H / 2.0 + 0.2], [0, 0, 1]], np.float32)
    return H, W, K, c2w


def make_image(H, W, seed):
    return np.random.RandomState(1000 + seed).uniform(0, 1, (H, W, 3)).astype(np.float32)


RAY_CASES = {
    'gl_center': dict(seed=1, H=12, W=16, inverse_y=False, flip_x=False, flip_y=False, mode='center', ndc=False),
    'cv_center': dict(seed=2, H=9, W=14, inverse_y=True, flip_x=False, flip_y=False, mode='center', ndc=False),
    'gl_flip_lefttop': dict(seed=3, H=10, W=7, inverse_y=False, flip_x=True, flip_y=True, mode='lefttop', ndc=False),
    'gl_ndc': dict(seed=4, H=8, W=8, inverse_y=False, flip_x=False, flip_y=False, mode='center', ndc=True),
}
TRAIN_VIEW_SIZES = [(40, 48), (33, 37), (40, 48)]
