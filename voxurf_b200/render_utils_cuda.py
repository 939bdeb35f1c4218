"""Drop-in for the reference's JIT module `render_utils_cuda`: same function names, argument order
and return lists as the pybind table /root/reference/lib/cuda/render_utils.cpp:170-184, implemented
on libvoxurf_b200.so.  Tensors in, tensors out; errors are RuntimeError like TORCH_CHECK."""
import torch

from ._lib import call


def _f32(t):
    return t if t.dtype == torch.float32 else t.float()


def infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far):
    n = rays_o.shape[0]
    t_min = torch.empty(n, dtype=torch.float32, device=rays_o.device)
    t_max = torch.empty_like(t_min)
    call('vx_infer_t_minmax', rays_o, rays_d, xyz_min, xyz_max, near, far, n, t_min, t_max)
    return [t_min, t_max]


def infer_n_samples(rays_d, t_min, t_max, stepdist):
    n = t_min.shape[0]
    out = torch.empty(n, dtype=torch.int64, device=t_min.device)
    call('vx_infer_n_samples', rays_d, t_min, t_max, stepdist, n, out)
    return out


def infer_ray_start_dir(rays_o, rays_d, t_min):
    n = rays_o.shape[0]
    start, dirs = torch.empty_like(rays_o), torch.empty_like(rays_o)
    call('vx_infer_ray_start_dir', rays_o, rays_d, t_min, n, start, dirs)
    return [start, dirs]


def ray_setup(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist):
    """Phase 1 of sample_pts_on_rays: everything per-ray + the exclusive scan of N_steps (one launch)."""
    n, dev = rays_o.shape[0], rays_o.device
    t_min = torch.empty(n, dtype=torch.float32, device=dev)
    t_max = torch.empty_like(t_min)
    n_steps = torch.empty(n, dtype=torch.int64, device=dev)
    start, dirs = torch.empty_like(rays_o), torch.empty_like(rays_o)
    offsets = torch.empty(n + 1, dtype=torch.int64, device=dev)
    call('vx_ray_setup', rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist, n, t_min, t_max, n_steps, start, dirs,
         offsets)
    return t_min, t_max, n_steps, start, dirs, offsets


def sample_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist):
    """-> [rays_pts (M,3), mask_outbbox (M,) bool, ray_id (M,) int64, step_id (M,) int64, N_steps, t_min, t_max]"""
    dev = rays_o.device
    n = rays_o.shape[0]
    t_min, t_max, n_steps, start, dirs, offsets = ray_setup(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist)
    total = int(offsets[n].item())  # the reference syncs here too (render_utils_kernel.cu:212)
    pts = torch.empty(total, 3, dtype=rays_o.dtype, device=dev)
    mask = torch.empty(total, dtype=torch.bool, device=dev)
    ray_id = torch.empty(total, dtype=torch.int64, device=dev)
    step_id = torch.empty(total, dtype=torch.int64, device=dev)
    if total > 0:
        call('vx_sample_fill', start, dirs, xyz_min, xyz_max, offsets, n, stepdist, pts, mask, ray_id, step_id)
    return [pts, mask, ray_id, step_id, n_steps, t_min, t_max]


def sample_ndc_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, N_samples):
    n, dev = rays_o.shape[0], rays_o.device
    pts = torch.empty(n, N_samples, 3, dtype=rays_o.dtype, device=dev)
    mask = torch.empty(n, N_samples, dtype=torch.bool, device=dev)
    call('vx_sample_ndc_pts_on_rays', rays_o, rays_d, xyz_min, xyz_max, N_samples, n, pts, mask)
    return [pts, mask]


def sample_bg_pts_on_rays(rays_o, rays_d, t_max, bg_preserve, N_samples):
    n = rays_o.shape[0]
    pts = torch.empty(n, N_samples, 3, dtype=rays_o.dtype, device=rays_o.device)
    call('vx_sample_bg_pts_on_rays', rays_o, rays_d, t_max, bg_preserve, N_samples, n, pts)
    return pts


def maskcache_lookup(world, xyz, xyz2ijk_scale, xyz2ijk_shift):
    n = xyz.shape[0]
    out = torch.zeros(n, dtype=torch.bool, device=xyz.device)
    if n == 0:
        return out
    call('vx_maskcache_lookup', world, xyz, xyz2ijk_scale, xyz2ijk_shift, world.shape[0], world.shape[1],
         world.shape[2], n, out)
    return out


def raw2alpha(density, shift, interval):
    exp_d, alpha = torch.empty_like(density), torch.empty_like(density)
    call('vx_raw2alpha', density, shift, None, interval, density.shape[0], exp_d, alpha)
    return [exp_d, alpha]


def raw2alpha_nonuni(density, shift, interval):
    exp_d, alpha = torch.empty_like(density), torch.empty_like(density)
    call('vx_raw2alpha', density, shift, interval, 0.0, density.shape[0], exp_d, alpha)
    return [exp_d, alpha]


def raw2alpha_backward(exp, grad_back, interval):
    grad = torch.empty_like(exp)
    call('vx_raw2alpha_backward', exp, grad_back, None, interval, exp.shape[0], grad)
    return grad


def raw2alpha_nonuni_backward(exp, grad_back, interval):
    grad = torch.empty_like(exp)
    call('vx_raw2alpha_backward', exp, grad_back, interval, 0.0, exp.shape[0], grad)
    return grad


def alpha2weight(alpha, ray_id, n_rays):
    """-> [weight, T, alphainv_last, i_start, i_end]; ray_id int64, sorted."""
    dev = alpha.device
    n = alpha.shape[0]
    weight, T = torch.empty_like(alpha), torch.empty_like(alpha)
    alphainv_last = torch.empty(n_rays, dtype=alpha.dtype, device=dev)
    i_start = torch.empty(n_rays, dtype=torch.int64, device=dev)
    i_end = torch.empty(n_rays, dtype=torch.int64, device=dev)
    call('vx_alpha2weight', alpha, ray_id, n, n_rays, weight, T, alphainv_last, i_start, i_end)
    return [weight, T, alphainv_last, i_start, i_end]


def alpha2weight_backward(alpha, weight, T, alphainv_last, i_start, i_end, n_rays, grad_weights, grad_last):
    grad = torch.empty_like(alpha)
    call('vx_alpha2weight_backward', alpha, weight, T, alphainv_last, i_start, i_end, n_rays, alpha.shape[0],
         grad_weights.contiguous(), grad_last.contiguous(), grad)
    return grad
