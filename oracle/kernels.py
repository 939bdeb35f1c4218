"""TEST INFRASTRUCTURE ONLY -- CPU oracle.  Never imported by voxurf_b200/.

ctypes front-end of oracle/ref_kernels.c (the plain-C restatement of the reference's native
operators).  Functions take and return CPU torch tensors with the dtypes and shapes of the
reference's pybind surface (/root/reference/lib/cuda/render_utils.cpp:170-184,
total_variation.cpp:29-32, adam_upd.cpp:79-86).
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, 'ref_kernels.c')
_SO = os.path.join(_HERE, 'libvoxurf_oracle.so')
_lib = None


def build(force=False):
    """gcc -O2 -ffp-contract=off: no implicit FMA; the source spells the fusions nvcc makes."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-fvisibility=hidden', '-shared',
                               '-fPIC', '-o', _SO, _SRC, '-lm'])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _f(x):
    return ctypes.c_float(float(x))


def _f32(t):
    return t.detach().to(torch.float32).contiguous()


def infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far):
    rays_o, rays_d, xyz_min, xyz_max = map(_f32, (rays_o, rays_d, xyz_min, xyz_max))
    n = rays_o.shape[0]
    t_min, t_max = torch.empty(n), torch.empty(n)
    lib().vxo_infer_t_minmax(_p(rays_o), _p(rays_d), _p(xyz_min), _p(xyz_max), _f(near), _f(far),
                             ctypes.c_int(n), _p(t_min), _p(t_max))
    return t_min, t_max


def infer_n_samples(rays_d, t_min, t_max, stepdist):
    rays_d, t_min, t_max = map(_f32, (rays_d, t_min, t_max))
    n = t_min.shape[0]
    out = torch.empty(n, dtype=torch.int64)
    lib().vxo_infer_n_samples(_p(rays_d), _p(t_min), _p(t_max), _f(stepdist), ctypes.c_int(n), _p(out))
    return out


def infer_ray_start_dir(rays_o, rays_d, t_min):
    rays_o, rays_d, t_min = map(_f32, (rays_o, rays_d, t_min))
    n = rays_o.shape[0]
    start, dirs = torch.empty(n, 3), torch.empty(n, 3)
    lib().vxo_infer_ray_start_dir(_p(rays_o), _p(rays_d), _p(t_min), ctypes.c_int(n), _p(start), _p(dirs))
    return start, dirs


def sample_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist):
    """render_utils_kernel.cu:196-242 -> [pts, mask_outbbox, ray_id, step_id, N_steps, t_min, t_max]"""
    xyz_min, xyz_max = _f32(xyz_min), _f32(xyz_max)
    t_min, t_max = infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far)
    n_steps = infer_n_samples(rays_d, t_min, t_max, stepdist)
    start, dirs = infer_ray_start_dir(rays_o, rays_d, t_min)
    total = int(n_steps.sum())
    pts = torch.empty(total, 3)
    mask = torch.empty(total, dtype=torch.bool)
    ray_id = torch.empty(total, dtype=torch.int64)
    step_id = torch.empty(total, dtype=torch.int64)
    lib().vxo_sample_pts_fill(_p(start), _p(dirs), _p(xyz_min), _p(xyz_max), _p(n_steps),
                              ctypes.c_int(rays_o.shape[0]), _f(stepdist), _p(pts), _p(mask),
                              _p(ray_id), _p(step_id))
    return pts, mask, ray_id, step_id, n_steps, t_min, t_max


def sample_ndc_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, n_samples):
    rays_o, rays_d, xyz_min, xyz_max = map(_f32, (rays_o, rays_d, xyz_min, xyz_max))
    n = rays_o.shape[0]
    pts = torch.empty(n, n_samples, 3)
    mask = torch.empty(n, n_samples, dtype=torch.bool)
    lib().vxo_sample_ndc_pts(_p(rays_o), _p(rays_d), _p(xyz_min), _p(xyz_max), ctypes.c_int(n_samples),
                             ctypes.c_int(n), _p(pts), _p(mask))
    return pts, mask


def sample_bg_pts_on_rays(rays_o, rays_d, t_max, bg_preserve, n_samples):
    rays_o, rays_d, t_max = map(_f32, (rays_o, rays_d, t_max))
    n = rays_o.shape[0]
    pts = torch.empty(n, n_samples, 3)
    lib().vxo_sample_bg_pts(_p(rays_o), _p(rays_d), _p(t_max), _f(bg_preserve), ctypes.c_int(n_samples),
                            ctypes.c_int(n), _p(pts))
    return pts


def maskcache_lookup(world, xyz, scale, shift):
    world = world.to(torch.bool).contiguous()
    xyz, scale, shift = map(_f32, (xyz, scale, shift))
    n = xyz.shape[0]
    out = torch.empty(n, dtype=torch.bool)
    lib().vxo_maskcache_lookup(_p(world), _p(xyz), _p(scale), _p(shift), ctypes.c_int(world.shape[0]),
                               ctypes.c_int(world.shape[1]), ctypes.c_int(world.shape[2]),
                               ctypes.c_int64(n), _p(out))
    return out


def raw2alpha(density, shift, interval):
    density = _f32(density)
    iv = _f32(interval) if torch.is_tensor(interval) else None
    e, a = torch.empty_like(density), torch.empty_like(density)
    lib().vxo_raw2alpha(_p(density), _f(shift), _p(iv), _f(0.0 if iv is not None else interval),
                        ctypes.c_int64(density.numel()), _p(e), _p(a))
    return e, a


def raw2alpha_backward(exp_d, grad_back, interval):
    exp_d, grad_back = _f32(exp_d), _f32(grad_back)
    iv = _f32(interval) if torch.is_tensor(interval) else None
    g = torch.empty_like(exp_d)
    lib().vxo_raw2alpha_backward(_p(exp_d), _p(grad_back), _p(iv), _f(0.0 if iv is not None else interval),
                                 ctypes.c_int64(exp_d.numel()), _p(g))
    return g


def alpha2weight(alpha, ray_id, n_rays):
    alpha = _f32(alpha)
    ray_id = ray_id.to(torch.int64).contiguous()
    n = alpha.shape[0]
    w, T = torch.empty(n), torch.empty(n)
    last = torch.empty(n_rays)
    i_start = torch.empty(n_rays, dtype=torch.int64)
    i_end = torch.empty(n_rays, dtype=torch.int64)
    lib().vxo_alpha2weight(_p(alpha), _p(ray_id), ctypes.c_int64(n), ctypes.c_int(n_rays), _p(w), _p(T),
                           _p(last), _p(i_start), _p(i_end))
    return w, T, last, i_start, i_end


def alpha2weight_backward(alpha, weight, T, alphainv_last, i_start, i_end, n_rays, grad_weights, grad_last):
    alpha, weight, T, alphainv_last, grad_weights, grad_last = map(
        _f32, (alpha, weight, T, alphainv_last, grad_weights, grad_last))
    g = torch.empty_like(alpha)
    lib().vxo_alpha2weight_backward(_p(alpha), _p(weight), _p(T), _p(alphainv_last), _p(i_start.contiguous()),
                                    _p(i_end.contiguous()), ctypes.c_int(n_rays), ctypes.c_int64(alpha.numel()),
                                    _p(grad_weights), _p(grad_last), _p(g))
    return g


def total_variation_add_grad(param, grad, wx, wy, wz, dense_mode, mask=None):
    """In place on `grad` (total_variation.cpp:13-27). param/grad: (1,C,X,Y,Z) contiguous fp32."""
    assert param.is_contiguous() and grad.is_contiguous() and param.dtype == torch.float32
    m = _f32(mask) if mask is not None else None
    lib().vxo_total_variation_add_grad(_p(param), _p(grad), _p(m), _f(wx), _f(wy), _f(wz),
                                       ctypes.c_int(int(dense_mode)), ctypes.c_int64(param.shape[2]),
                                       ctypes.c_int64(param.shape[3]), ctypes.c_int64(param.shape[4]),
                                       ctypes.c_int64(param.numel()))


def adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps, mode=0, perlr=None):
    """In place (adam_upd.cpp:36-77). mode 0 dense, 1 masked (grad != 0), 2 per-voxel lr."""
    for t in (param, grad, exp_avg, exp_avg_sq):
        assert t.is_contiguous() and t.dtype == torch.float32
    lib().vxo_adam_upd(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), _p(perlr), ctypes.c_int64(param.numel()),
                       ctypes.c_int(step), _f(beta1), _f(beta2), _f(lr), _f(eps), ctypes.c_int(mode))


def cumdist_thres(dist, thres):
    """ub360_utils_kernel.cu:13-47 -> bool (n_rays, n_pts)"""
    dist = dist.contiguous().float()
    mask = torch.zeros(dist.shape, dtype=torch.uint8)
    lib().vxo_cumdist_thres(_p(dist), ctypes.c_float(thres), ctypes.c_int(dist.shape[0]), ctypes.c_int(dist.shape[1]), _p(mask))
    return mask.bool()
