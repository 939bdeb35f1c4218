"""Shared host-side machinery of the two model mirrors (voxurf_fine.Voxurf, voxurf_coarse.Voxurf): grid
geometry, the fused ray march, the free-space mask bookkeeping and the regulariser hooks.  The reference
duplicates this code in lib/voxurf_fine.py and lib/voxurf_coarse.py; citations below name the fine file, the
coarse file has the same functions at the lines given in each subclass."""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import grid, ops
from ._lib import call
from .grid import MaskCache


def _gaussian_weights(ksize, sigma):
    """lib/voxurf_fine.py:246-254 -> flat python list (k^3,), float32-rounded like the reference's Conv3d weight."""
    r = np.arange(-(ksize // 2), ksize // 2 + 1, 1)
    xx, yy, zz = np.meshgrid(r, r, r)
    k = torch.from_numpy(np.exp(-(xx ** 2 + yy ** 2 + zz ** 2) / (2 * sigma ** 2))).float()
    return (k / k.sum()).flatten().tolist()


def _binomial_weights():
    """tv_smooth_conv weights, lib/voxurf_fine.py:208-239 with sigma = 0."""
    k = np.asarray([[[1, 2, 1], [2, 4, 2], [1, 2, 1]], [[2, 4, 2], [4, 8, 4], [2, 4, 2]], [[1, 2, 1], [2, 4, 2], [1, 2, 1]]],
                   dtype=np.float64)
    return torch.from_numpy(k / k.sum()).float().flatten().tolist()


class FrozenConv(nn.Module):
    """Holder of the `weight` / `bias` tensors of a frozen nn.Conv3d of the reference (grad_conv, tv_smooth_conv,
    smooth_conv*: lib/voxurf_fine.py:204-258) so that state_dict keys and shapes match the reference's checkpoints
    (`<name>.weight`, `<name>.bias`).  The arithmetic runs in the stencil kernels from host-side weight lists."""

    def __init__(self, weight, bias_len):
        super().__init__()
        self.weight = nn.Parameter(weight.float().clone(), requires_grad=False)
        self.bias = nn.Parameter(torch.zeros(bias_len), requires_grad=False)


class SmoothConv(FrozenConv):
    """Stand-in for the frozen nn.Conv3d the reference builds in _gaussian_3dconv (lib/voxurf_fine.py:246-258)."""

    def __init__(self, ksize, sigma):
        weight_host = _gaussian_weights(ksize, sigma)
        super().__init__(torch.tensor(weight_host).view(1, 1, ksize, ksize, ksize), 1)
        self.ksize, self.sigma, self.weight_host = ksize, sigma, weight_host
        # the normalised Gaussian factorises: k^3 weights = w1 (x) w1 (x) w1 -> three 1-D passes instead of k^3 taps
        r = np.arange(-(ksize // 2), ksize // 2 + 1, 1)
        g = np.exp(-(r.astype(np.float64) ** 2) / (2 * sigma ** 2))
        self.weight1d_host = [float(np.float32(v)) for v in g / g.sum()]

    def forward(self, x):
        return ops.conv3d_replicate(x, self.weight_host, self.ksize, weight1d=self.weight1d_host)


def _grad_conv_weight(voxel_size):
    """lib/voxurf_fine.py:204-227 with sigma = 0 (state_dict compatibility; the 'interpolate' gradient mode never uses it)"""
    kernel = np.asarray([[[1, 2, 1], [2, 4, 2], [1, 2, 1]], [[2, 4, 2], [4, 8, 4], [2, 4, 2]], [[1, 2, 1], [2, 4, 2], [1, 2, 1]]],
                        dtype=np.float64)
    kernel1 = kernel / (kernel[0].sum() * 2 * voxel_size)
    weight = torch.from_numpy(np.concatenate([kernel1[None] for _ in range(3)])).float()
    weight[0, 1, :, :] *= 0
    weight[0, 0, :, :] *= -1
    weight[1, :, 1, :] *= 0
    weight[1, :, 0, :] *= -1
    weight[2, :, :, 1] *= 0
    weight[2, :, :, 0] *= -1
    return weight.unsqueeze(1).float()


def _mlp(dim0, width, depth):
    return nn.Sequential(
        nn.Linear(dim0, width), nn.ReLU(inplace=True),
        *[nn.Sequential(nn.Linear(width, width), nn.ReLU(inplace=True)) for _ in range(depth - 2)],
        nn.Linear(width, 3))



class VoxurfBase(nn.Module):
    def _init_common(self, xyz_min, xyz_max, num_voxels, num_voxels_base, alpha_init, s_ratio, s_start, s_learn,
                     step_start, fast_color_thres, nearest):
        """lib/voxurf_fine.py:51-86 (buffers, s_val, base voxel size, density + ball-initialised SDF grid)."""
        self.register_buffer('xyz_min', torch.Tensor(list(xyz_min)))
        self.register_buffer('xyz_max', torch.Tensor(list(xyz_max)))
        self._min_host = [float(v) for v in xyz_min]
        self._max_host = [float(v) for v in xyz_max]
        self.fast_color_thres = fast_color_thres
        self.nearest = nearest
        self.s_ratio, self.s_start, self.s_learn, self.step_start = s_ratio, s_start, s_learn, step_start
        self.s_val = nn.Parameter(torch.ones(1) * s_start, requires_grad=False)
        self.sdf_init_mode = 'ball_init'
        self.num_voxels_base = num_voxels_base
        self.voxel_size_base = ((self.xyz_max - self.xyz_min).prod() / self.num_voxels_base).pow(1 / 3)
        self.alpha_init = alpha_init
        self.act_shift = np.log(1 / (1 - alpha_init) - 1)
        self._set_grid_resolution(num_voxels)
        self.density = nn.Parameter(torch.zeros([1, 1, *self.world_size]))
        self.sdf = grid.create_grid('DenseGrid', channels=1, world_size=self.world_size, xyz_min=self.xyz_min,
                                    xyz_max=self.xyz_max)
        ws = [int(w) for w in self.world_size]
        x, y, z = np.mgrid[-1.0:1.0:ws[0] * 1j, -1.0:1.0:ws[1] * 1j, -1.0:1.0:ws[2] * 1j]
        self.sdf.grid.data = torch.from_numpy((x ** 2 + y ** 2 + z ** 2) ** 0.5 - 1).float()[None, None, ...]
        self.register_buffer('nonempty_mask', None)   # becomes a real buffer in _set_nonempty_mask (lib/voxurf_fine.py:362-365)
        self._tv_smooth_w = _binomial_weights()
        # frozen convs of the reference (lib/voxurf_fine.py:204-239): kept for their state_dict keys
        self.grad_conv = FrozenConv(_grad_conv_weight(self._voxel_size_host), 3)
        self.tv_smooth_conv = FrozenConv(torch.tensor(self._tv_smooth_w).view(1, 1, 3, 3, 3), 1)
        self.gradient = None

    def _init_mask_cache(self, mask_cache_path, mask_cache_thres, mask_cache_state):
        """lib/voxurf_fine.py:189-199.  The nonempty mask is built by _set_nonempty_mask() once the model is on
        its device (the reference does it in the constructor because its default tensors are CUDA, run.py:954)."""
        self.mask_cache_path, self.mask_cache_thres = mask_cache_path, mask_cache_thres
        if mask_cache_state is not None or (mask_cache_path is not None and mask_cache_path):
            self.mask_cache = MaskCache(path=mask_cache_path, mask_cache_thres=mask_cache_thres, state=mask_cache_state)
        else:
            self.mask_cache = None

    def _update_s_val(self, global_step):
        """lib/voxurf_fine.py:466-473 -> (s_val to report, 1/s_val as float32)."""
        if global_step is not None:
            s_val = 1. / (global_step + self.s_ratio / self.s_start - self.step_start) * self.s_ratio
            self.s_val.data.fill_(s_val)       # (== torch.ones_like(self.s_val) * s_val of the reference, one launch)
            self._s_val_host = float(np.float32(s_val))
        else:
            s_val = 0
            if not hasattr(self, '_s_val_host'):
                self._s_val_host = float(self.s_val.item())
        return s_val, float(np.float32(1.0) / np.float32(self._s_val_host))

    def _set_grid_resolution(self, num_voxels):
        """lib/voxurf_fine.py:315-324"""
        self.num_voxels = num_voxels
        self.voxel_size = ((self.xyz_max - self.xyz_min).prod() / num_voxels).pow(1 / 3)
        self.world_size = ((self.xyz_max - self.xyz_min) / self.voxel_size).long()
        self.voxel_size_ratio = self.voxel_size / self.voxel_size_base
        self._voxel_size_host = float(self.voxel_size)

    def init_smooth_conv(self, ksize=3, sigma=1):
        """lib/voxurf_fine.py:268-272"""
        self.smooth_sdf = ksize > 0
        if self.smooth_sdf:
            self.smooth_conv = SmoothConv(ksize, sigma)

    def _gaussian_3dconv(self, ksize=3, sigma=1):
        return SmoothConv(ksize, sigma)

    def get_MaskCache_kwargs(self):
        """lib/voxurf_fine.py:344-351"""
        return {'xyz_min': self.xyz_min.cpu().numpy(), 'xyz_max': self.xyz_max.cpu().numpy(),
                'act_shift': self.act_shift, 'voxel_size_ratio': self.voxel_size_ratio, 'nearest': self.nearest}

    @torch.no_grad()
    def _set_nonempty_mask(self):
        """lib/voxurf_fine.py:353-367: mask-cache query on the grid lattice; empty voxels get sdf = 1."""
        dev = self.sdf.grid.device
        ws = self.density.shape[2:]
        xyz = torch.stack(torch.meshgrid(
            torch.linspace(self._min_host[0], self._max_host[0], ws[0]),
            torch.linspace(self._min_host[1], self._max_host[1], ws[1]),
            torch.linspace(self._min_host[2], self._max_host[2], ws[2]), indexing='ij'), -1).to(dev)
        self.nonempty_mask = self.mask_cache(xyz)[None, None].contiguous()
        self._refresh_derived()
        self.density[~self.nonempty_mask] = -100
        self.sdf.grid[~self.nonempty_mask] = 1

    def _refresh_derived(self):
        """Host-side values derived from buffers / parameters: recomputed whenever those change underneath us."""
        self._n_nonempty = int(self.nonempty_mask.sum().item()) if self.nonempty_mask is not None else 0
        self._s_val_host = float(self.s_val.detach().reshape(-1)[0].item())

    def load_state_dict(self, state_dict, strict=True, **kw):
        """Reference checkpoints (run.py:786-793, lib/utils.py:232-300) load with strict=True: same keys and shapes.
        A checkpoint taken after the non-empty mask was built carries a `nonempty_mask` the fresh model may not have."""
        if 'nonempty_mask' in state_dict and self.nonempty_mask is None:
            self.nonempty_mask = torch.zeros_like(state_dict['nonempty_mask'], device=self.sdf.grid.device)
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._refresh_derived()
        return out

    @torch.no_grad()
    def scale_volume_grid(self, num_voxels):
        """lib/voxurf_fine.py:384-397"""
        self._set_grid_resolution(num_voxels)
        ws = tuple(int(w) for w in self.world_size)
        self.density = nn.Parameter(F.interpolate(self.density.data, size=ws, mode='trilinear', align_corners=True))
        self.sdf.scale_volume_grid(self.world_size)
        self.k0.scale_volume_grid(self.world_size)
        if self.mask_cache is not None:
            self._set_nonempty_mask()

    def sdf_total_variation_add_grad(self, weight, dense_mode):
        """lib/voxurf_fine.py:403-405"""
        w = weight * int(self.world_size.max()) / 128
        self.sdf.total_variation_add_grad(w, w, w, dense_mode)

    def k0_total_variation_add_grad(self, weight, dense_mode):
        """lib/voxurf_fine.py:407-409"""
        w = weight * int(self.world_size.max()) / 128
        self.k0.total_variation_add_grad(w, w, w, dense_mode)

    def neus_sdf_gradient(self, mode=None, sdf=None):
        """lib/voxurf_fine.py:440-460 ('interpolate')"""
        if sdf is None:
            sdf = self.sdf.grid
        return ops.fd_gradient(sdf, self._voxel_size_host)

    def _max_steps(self, stepdist):
        diag = math.sqrt(sum((b - a) ** 2 for a, b in zip(self._min_host, self._max_host)))
        return int(math.ceil(diag / stepdist)) + 2

    def _march(self, rays_o, rays_d, near, stepsize, want_mask_outbbox=True):
        """Fused lib/voxurf_fine.py:593-617 + :631-636.  -> dict with int32 ray_id/step_id (M2,), start, dirs, ..."""
        from . import render_utils_cuda as ru
        far = 1e9
        rays_o, rays_d = rays_o.contiguous(), rays_d.contiguous()
        N, dev = rays_o.shape[0], rays_o.device
        stepdist = float(stepsize) * self._voxel_size_host
        stepdist = float(np.float32(stepdist))
        t_min, t_max, n_steps, start, dirs, offsets = ru.ray_setup(rays_o, rays_d, self.xyz_min, self.xyz_max, near, far, stepdist)
        words = N * (self._max_steps(stepdist) // 32 + 2) + 1
        bits_in = torch.empty(words, dtype=torch.int32, device=dev)
        bits_keep = torch.empty(words, dtype=torch.int32, device=dev)
        keep_count = torch.empty(N, dtype=torch.int32, device=dev)
        keep_off = torch.empty(N + 1, dtype=torch.int32, device=dev)
        mc = self.mask_cache.march_args() if self.mask_cache is not None else (None, 1, 1, 1, [0., 0., 0.], [1., 1., 1.], 0., 1., 0.)
        call('vx_march_flags', start, dirs, self.xyz_min, self.xyz_max, offsets, N, stepdist, *mc, bits_in, bits_keep,
             keep_count, keep_off)
        totals = torch.stack([offsets[N], keep_off[N].to(torch.int64)]).cpu()   # the ONE sync of sampling
        M0, M2 = int(totals[0]), int(totals[1])
        ray_id = torch.empty(M2, dtype=torch.int32, device=dev)
        step_id = torch.empty(M2, dtype=torch.int32, device=dev)
        mask_outbbox = torch.empty(M0, dtype=torch.bool, device=dev) if want_mask_outbbox else None
        call('vx_march_emit', offsets, N, bits_keep, keep_off, M2, ray_id, step_id, mask_outbbox)
        pts = torch.empty(M2, 3, dtype=torch.float32, device=dev)
        call('vx_points_from_steps', ray_id, step_id, start, dirs, stepdist, None, M2, pts)
        return dict(ray_pts=pts, ray_id=ray_id, step_id=step_id, mask_outbbox=mask_outbbox, n_steps=n_steps,
                    keep_off=keep_off, start=start, dirs=dirs, stepdist=stepdist, t_min=t_min, t_max=t_max)

    def sample_ray(self, rays_o, rays_d, near, far, stepsize, **render_kwargs):
        """lib/voxurf_fine.py:593-617 (legacy form: in-bbox samples only, int64 ids)."""
        from . import render_utils_cuda as ru
        stepdist = float(np.float32(float(stepsize) * self._voxel_size_host))
        ray_pts, mask_outbbox, ray_id, step_id, N_steps, t_min, t_max = ru.sample_pts_on_rays(
            rays_o.contiguous(), rays_d.contiguous(), self.xyz_min, self.xyz_max, near, 1e9, stepdist)
        N_steps = ray_id.unique(return_counts=True)[1]
        inb = ~mask_outbbox
        return ray_pts[inb], ray_id[inb], step_id[inb], mask_outbbox, N_steps

    def hit_coarse_geo(self, rays_o, rays_d, near, far, stepsize, **render_kwargs):
        """lib/voxurf_fine.py:579-591: which rays have at least one sample inside the mask cache."""
        shape = rays_o.shape[:-1]
        o, d = rays_o.reshape(-1, 3).contiguous(), rays_d.reshape(-1, 3).contiguous()
        hit = torch.empty(o.shape[0], dtype=torch.bool, device=o.device)
        stepdist = float(np.float32(float(stepsize) * self._voxel_size_host))
        # warp-per-ray march that stops at the first sample inside the mask: nothing per-sample is materialised
        # (the reference's sample_pts_on_rays + MaskCache on every sample is why it feeds this 64 image rows at a time)
        call('vx_rays_hit_mask', o, d, o.shape[0], self._min_host, self._max_host, float(near), 1e9, stepdist,
             *self.mask_cache.march_args(), hit)
        return hit.reshape(shape)

