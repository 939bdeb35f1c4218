"""Host-side mirror of the reference's fine-stage model, lib/voxurf_fine.py (class Voxurf, MaskCache,
Alphas2Weights), running on the B200 operators of libvoxurf_b200.so.

Drop-in scope: same constructor keywords, attribute names, state_dict keys, `forward()` signature and
`ret_dict` keys as lib/voxurf_fine.py:25-203,620-802, so the reference's train / render loops
(run.py:552-671, :81-227) work on it unchanged.  What differs is *how* the step runs on the GPU:

  reference                                              here
  ---------                                              ----
  sample_pts_on_rays (8 launches, 145 MB of M0-sized     one ray-setup launch + warp-per-ray keep-bit pass +
  lists) + 2 boolean-mask compactions + MaskCache        emit pass; only the compact int32 (ray_id, step_id) list
  grid_sample + 5 elementwise                            is written
  7 x F.grid_sample + ~15 elementwise for sdf/gradient   one fused tap kernel (vx_sdf_taps, L=1)
  ~20 elementwise for the NeuS alpha                     one kernel (vx_neus_alpha), one for its backward
  thread-per-ray alpha2weight (+ host-indexed write)     warp-per-ray lock-step recurrence, bit-exact, no sync
  24 + C x F.grid_sample at the MLP rows                 vx_sdf_taps (L=4, normalised) + vx_grid_gather
  ATen grid_sampler backward (atomics incl. zeros)       per-sample combined scatter that skips exact zeros
  torch_scatter.segment_coo                              deterministic segmented sum
"""
import math
from copy import deepcopy

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import grid, ops
from ._lib import call
from .grid import MaskCache  # noqa: F401  (lib/voxurf_fine.py:917 lives next to the model in the reference)
from .ops import Alphas2Weights  # noqa: F401
from .torch_scatter import segment_coo


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


class SmoothConv:
    """Callable stand-in for the frozen nn.Conv3d the reference builds in _gaussian_3dconv."""

    def __init__(self, ksize, sigma):
        self.ksize, self.sigma = ksize, sigma
        self.weight_host = _gaussian_weights(ksize, sigma)

    def __call__(self, x):
        return ops.conv3d_replicate(x, self.weight_host, self.ksize)


def _mlp(dim0, width, depth):
    return nn.Sequential(
        nn.Linear(dim0, width), nn.ReLU(inplace=True),
        *[nn.Sequential(nn.Linear(width, width), nn.ReLU(inplace=True)) for _ in range(depth - 2)],
        nn.Linear(width, 3))


class Voxurf(nn.Module):
    def __init__(self, xyz_min, xyz_max, num_voxels=0, num_voxels_base=0, alpha_init=None, nearest=False,
                 mask_cache_path=None, mask_cache_thres=1e-3, fast_color_thres=0, rgbnet_dim=0, rgbnet_direct=False,
                 rgbnet_full_implicit=False, rgbnet_depth=3, rgbnet_width=128, posbase_pe=5, viewbase_pe=4,
                 center_sdf=False, grad_feat=(1.0,), sdf_feat=(), use_layer_norm=False, grad_mode='interpolate',
                 s_ratio=2000, s_start=0.2, s_learn=False, step_start=0, smooth_sdf=False, smooth_ksize=0,
                 smooth_sigma=1, k_rgbnet_depth=3, k_res=False, k_posbase_pe=5, k_viewbase_pe=4, k_center_sdf=False,
                 k_grad_feat=(1.0,), k_sdf_feat=(), smooth_scale=True, use_grad_norm=True, use_rgb_k=True,
                 k_detach_1=True, k_detach_2=True, use_rgbnet_k0=False, mask_cache_state=None, k0_channels_last=False,
                 **kwargs):
        super().__init__()
        if nearest or use_layer_norm or s_learn or rgbnet_dim <= 0 or not use_rgb_k or use_rgbnet_k0 \
                or not (k_detach_1 and k_detach_2) or grad_mode != 'interpolate':
            raise NotImplementedError('option outside the configurations the reference ships (configs/*/fine.py)')
        self.register_buffer('xyz_min', torch.Tensor(list(xyz_min)))
        self.register_buffer('xyz_max', torch.Tensor(list(xyz_max)))
        self._min_host = [float(v) for v in xyz_min]
        self._max_host = [float(v) for v in xyz_max]
        self.fast_color_thres = fast_color_thres
        self.nearest = nearest
        self.smooth_scale = smooth_scale
        self.s_ratio, self.s_start, self.s_learn, self.step_start = s_ratio, s_start, s_learn, step_start
        self.s_val = nn.Parameter(torch.ones(1) * s_start, requires_grad=False)   # lib/voxurf_fine.py:61-62
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
        self.init_smooth_conv(smooth_ksize, smooth_sigma)
        self.rgbnet_kwargs = {'rgbnet_dim': rgbnet_dim, 'rgbnet_direct': rgbnet_direct,
                              'rgbnet_full_implicit': rgbnet_full_implicit, 'rgbnet_depth': rgbnet_depth,
                              'rgbnet_width': rgbnet_width, 'posbase_pe': posbase_pe, 'viewbase_pe': viewbase_pe}
        self.k0_dim = rgbnet_dim
        self.k0 = grid.create_grid('DenseGrid', channels=self.k0_dim, world_size=self.world_size, xyz_min=self.xyz_min,
                                   xyz_max=self.xyz_max, channels_last=k0_channels_last)
        self.rgbnet_direct = rgbnet_direct
        self.register_buffer('posfreq', torch.FloatTensor([(2 ** i) for i in range(posbase_pe)]))
        self.register_buffer('viewfreq', torch.FloatTensor([(2 ** i) for i in range(viewbase_pe)]))
        dim0 = (3 + 3 * posbase_pe * 2) + (3 + 3 * viewbase_pe * 2)
        self.use_grad_norm, self.center_sdf = use_grad_norm, center_sdf
        self.grad_feat, self.sdf_feat = tuple(grad_feat), tuple(sdf_feat)
        self.use_rgb_k, self.k_detach_1, self.k_detach_2 = use_rgb_k, k_detach_1, k_detach_2
        self.use_rgbnet_k0, self.use_layer_norm = use_rgbnet_k0, use_layer_norm
        dim0 += len(self.grad_feat) * 3 + len(self.sdf_feat) * 6 + (1 if center_sdf else 0)
        self.rgbnet = _mlp(dim0, rgbnet_width, rgbnet_depth)
        nn.init.constant_(self.rgbnet[-1].bias, 0)
        self.k_res, self.k_center_sdf = k_res, k_center_sdf
        self.k_grad_feat, self.k_sdf_feat = tuple(k_grad_feat), tuple(k_sdf_feat)
        self.register_buffer('k_posfreq', torch.FloatTensor([(2 ** i) for i in range(k_posbase_pe)]))
        self.register_buffer('k_viewfreq', torch.FloatTensor([(2 ** i) for i in range(k_viewbase_pe)]))
        k_dim0 = (3 + 3 * k_posbase_pe * 2) + (3 + 3 * k_viewbase_pe * 2) + self.k0_dim
        k_dim0 += (3 if k_res else 0) + (1 if k_center_sdf else 0) + len(self.k_grad_feat) * 3 + len(self.k_sdf_feat) * 6
        self.k_rgbnet = _mlp(k_dim0, rgbnet_width, k_rgbnet_depth)
        # (the reference zeroes rgbnet[-1].bias twice and leaves k_rgbnet's at its default, lib/voxurf_fine.py:185)
        self.mask_cache_path, self.mask_cache_thres = mask_cache_path, mask_cache_thres
        if mask_cache_state is not None or (mask_cache_path is not None and mask_cache_path):
            self.mask_cache = MaskCache(path=mask_cache_path, mask_cache_thres=mask_cache_thres, state=mask_cache_state)
        else:
            self.mask_cache = None
        self.nonempty_mask = None
        self.grad_mode = grad_mode
        self._tv_smooth_w = _binomial_weights()
        self.gradient = None

    # ------------------------------------------------------------------ construction helpers
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

    def get_kwargs(self):
        """lib/voxurf_fine.py:326-342"""
        return {'xyz_min': self.xyz_min.cpu().numpy(), 'xyz_max': self.xyz_max.cpu().numpy(),
                'num_voxels': self.num_voxels, 'num_voxels_base': self.num_voxels_base, 'alpha_init': self.alpha_init,
                'nearest': self.nearest, 'mask_cache_path': self.mask_cache_path,
                'mask_cache_thres': self.mask_cache_thres, 'fast_color_thres': self.fast_color_thres,
                'grad_feat': self.grad_feat, 'sdf_feat': self.sdf_feat, 'k_grad_feat': self.k_grad_feat,
                'k_sdf_feat': self.k_sdf_feat, **self.rgbnet_kwargs}

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
        self._n_nonempty = int(self.nonempty_mask.sum().item())
        self.density[~self.nonempty_mask] = -100
        self.sdf.grid[~self.nonempty_mask] = 1

    def init_sdf_from_sdf(self, sdf0=None, smooth=False, reduce=1., ksize=3, sigma=1., zero2neg=True):
        """lib/voxurf_fine.py:280-296 (coarse -> fine hand-off; once, off the hot path)."""
        with torch.no_grad():
            if sdf0.shape != self.sdf.grid.shape:
                sdf0 = F.interpolate(sdf0, size=tuple(int(w) for w in self.world_size), mode='trilinear', align_corners=True)
            if smooth:
                sdf_data = SmoothConv(ksize, sigma)(sdf0.to(self.sdf.grid) / reduce)
                self.sdf.grid = nn.Parameter(sdf_data / reduce)
            else:
                self.sdf.grid.data = sdf0.to(self.sdf.grid) / reduce
            if self.mask_cache is not None:
                self._set_nonempty_mask()
            if self.smooth_scale:
                self.sdf.grid = nn.Parameter(SmoothConv(5, 1)(self.sdf.grid.data))
            self.gradient = self.neus_sdf_gradient()

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

    # ------------------------------------------------------------------ regularisers
    def sdf_total_variation_add_grad(self, weight, dense_mode):
        """lib/voxurf_fine.py:403-405"""
        w = weight * int(self.world_size.max()) / 128
        self.sdf.total_variation_add_grad(w, w, w, dense_mode)

    def k0_total_variation_add_grad(self, weight, dense_mode):
        """lib/voxurf_fine.py:407-409"""
        w = weight * int(self.world_size.max()) / 128
        self.k0.total_variation_add_grad(w, w, w, dense_mode)

    def density_total_variation(self, sdf_tv=0, smooth_grad_tv=0, grad_tv=0, smooth_sdf_tv=0):
        """lib/voxurf_fine.py:412-421 (the smooth_grad_tv branch is the one the shipped configs use)."""
        tv = 0
        if sdf_tv > 0:
            raise NotImplementedError('autograd-form sdf_tv is a coarse-stage (ori_tv) option; see voxurf_coarse')
        if smooth_grad_tv > 0:
            tv = tv + ops.smooth_grad_tv(self.gradient, self.nonempty_mask[0, 0], self._tv_smooth_w, smooth_grad_tv,
                                         self._n_nonempty)
        return tv

    def neus_sdf_gradient(self, mode=None, sdf=None):
        """lib/voxurf_fine.py:440-460 ('interpolate')"""
        if sdf is None:
            sdf = self.sdf.grid
        return ops.fd_gradient(sdf, self._voxel_size_host)

    # ------------------------------------------------------------------ samplers
    def grid_sampler(self, xyz, *grids, mode=None, align_corners=True, sample_ret=True, sample_grad=False, displace=0.1,
                     smooth=False):
        """lib/voxurf_fine.py:502-534"""
        shape = xyz.shape[:-1]
        g = self.smooth_conv(grids[0]) if smooth else grids[0]
        if not sample_grad:
            out = ops.grid_gather(g, xyz, self._min_host, self._max_host).reshape(*shape, g.shape[1])
            return out.squeeze(-1)
        sdf, feat, grad = ops.sdf_taps(g, xyz, self._min_host, self._max_host, [1.0], self._voxel_size_host,
                                       use_grad_norm=False, xyz_order=True, want_sdf=True)
        outs = ([sdf.reshape(*shape)] if sample_ret else []) + [grad, feat]
        return outs

    def sample_sdfs(self, xyz, *grids, displace_list, mode='bilinear', align_corners=True, use_grad_norm=False):
        """lib/voxurf_fine.py:537-577 -> feat (P, 6L), grad (P, 3L)"""
        return ops.sdf_taps(grids[0], xyz, self._min_host, self._max_host, list(displace_list), self._voxel_size_host,
                            use_grad_norm=use_grad_norm, xyz_order=False, want_sdf=False)

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
        m = self._march(rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), near, stepsize, want_mask_outbbox=False)
        cnt = m['keep_off'][1:] - m['keep_off'][:-1]
        return (cnt > 0).reshape(shape)

    # ------------------------------------------------------------------ forward
    def forward(self, rays_o, rays_d, viewdirs, global_step=None, **render_kwargs):
        """Volume rendering, lib/voxurf_fine.py:620-802."""
        ret_dict = {}
        N = len(rays_o)
        viewdirs = viewdirs.contiguous()
        m = self._march(rays_o, rays_d, render_kwargs['near'], render_kwargs['stepsize'])
        ray_pts, step_id, mask_outbbox = m['ray_pts'], m['step_id'].long(), m['mask_outbbox']
        ray_id = m['ray_id'].long()

        sdf_grid = self.smooth_conv(self.sdf.grid) if self.smooth_sdf else self.sdf.grid
        sdf, gradient, feat = self.grid_sampler(ray_pts, sdf_grid, sample_ret=True, sample_grad=True, displace=1.0)

        dist = render_kwargs['stepsize'] * self._voxel_size_host
        if global_step is not None:   # lib/voxurf_fine.py:466-469
            s_val = 1. / (global_step + self.s_ratio / self.s_start - self.step_start) * self.s_ratio
            self.s_val.data = torch.ones_like(self.s_val) * s_val
            self._s_val_host = float(np.float32(s_val))
        else:
            s_val = 0
            if not hasattr(self, '_s_val_host'):
                self._s_val_host = float(self.s_val.item())
        inv_s = float(np.float32(1.0) / np.float32(self._s_val_host))
        alpha = ops.neus_alpha(viewdirs, ray_id, sdf, gradient, float(np.float32(dist)), inv_s)

        mask = None
        if self.fast_color_thres > 0:   # :647-654
            mask = (alpha > self.fast_color_thres)
            alpha, ray_id, ray_pts, step_id, gradient, sdf = (t[mask] for t in (alpha, ray_id, ray_pts, step_id, gradient, sdf))
        weights, alphainv_last = Alphas2Weights.apply(alpha, ray_id, N)
        if self.fast_color_thres > 0:   # :668-676, weights are not recomputed
            mask = (weights > self.fast_color_thres)
            weights, alpha, ray_pts, ray_id, step_id, gradient, sdf = (
                t[mask] for t in (weights, alpha, ray_pts, ray_id, step_id, gradient, sdf))

        k0 = self.k0(ray_pts)
        all_grad_inds = sorted(set(self.grad_feat + self.k_grad_feat))
        all_sdf_inds = sorted(set(self.sdf_feat + self.k_sdf_feat))
        assert all_grad_inds == all_sdf_inds
        if len(all_grad_inds) > 0:
            all_feat, all_grad = self.sample_sdfs(ray_pts, sdf_grid, displace_list=deepcopy(all_grad_inds),
                                                  use_grad_norm=self.use_grad_norm)
        else:
            all_feat, all_grad = None, None

        if render_kwargs.get('materialize_gradient', True):   # :692 (every forward in the reference)
            self.gradient = self.neus_sdf_gradient()

        viewdirs_emb = (viewdirs.unsqueeze(-1) * self.viewfreq).flatten(-2)
        viewdirs_emb = torch.cat([viewdirs, viewdirs_emb.sin(), viewdirs_emb.cos()], -1)
        rays_xyz = (ray_pts - self.xyz_min) / (self.xyz_max - self.xyz_min)
        xyz_emb = (rays_xyz.unsqueeze(-1) * self.posfreq).flatten(-2)
        xyz_emb = torch.cat([rays_xyz, xyz_emb.sin(), xyz_emb.cos()], -1)
        rgb_feat = torch.cat([xyz_emb, viewdirs_emb.flatten(0, -2)[ray_id]], -1)
        hierarchical_feats = []
        if self.center_sdf:
            hierarchical_feats.append(sdf[:, None])
        if len(all_grad_inds) > 0:
            hierarchical_feats += [all_feat, all_grad]
        if len(hierarchical_feats) > 0:
            rgb_feat = torch.cat([rgb_feat, *hierarchical_feats], dim=-1)
        rgb_logit = self.rgbnet(rgb_feat)
        rgb = torch.sigmoid(rgb_logit)

        k_xyz_emb = (rays_xyz.unsqueeze(-1) * self.k_posfreq).flatten(-2)
        k_xyz_emb = torch.cat([rays_xyz, k_xyz_emb.sin(), k_xyz_emb.cos()], -1)
        k_viewdirs_emb = (viewdirs.unsqueeze(-1) * self.k_viewfreq).flatten(-2)
        k_viewdirs_emb = torch.cat([viewdirs, k_viewdirs_emb.sin(), k_viewdirs_emb.cos()], -1)
        k_rgb_feat = torch.cat([k0, k_xyz_emb, k_viewdirs_emb.flatten(0, -2)[ray_id]], -1)
        assert len(self.k_grad_feat) == 1 and self.k_grad_feat[0] == 1.0
        assert len(self.k_sdf_feat) == 0
        all_feats_ = [gradient]
        if self.k_center_sdf:
            all_feats_.append(sdf[:, None])
        k_rgb_feat = torch.cat([k_rgb_feat, *all_feats_], dim=-1)
        if self.k_res:
            k_rgb_feat = torch.cat([k_rgb_feat, rgb_logit.detach()], dim=-1)
        k_rgb_logit = rgb_logit.detach() + self.k_rgbnet(k_rgb_feat)
        k_rgb = torch.sigmoid(k_rgb_logit)
        bg = render_kwargs['bg']
        k_rgb_marched = segment_coo(src=(weights.unsqueeze(-1) * k_rgb), index=ray_id,
                                    out=torch.zeros([N, 3], device=rgb.device), reduce='sum') + alphainv_last[..., None] * bg
        k_rgb_marched = k_rgb_marched.clamp(0, 1)
        rgb_marched = segment_coo(src=(weights.unsqueeze(-1) * rgb), index=ray_id,
                                  out=torch.zeros([N, 3], device=rgb.device), reduce='sum') + alphainv_last[..., None] * bg

        if gradient is not None and render_kwargs.get('render_grad', False):
            normal = gradient / (gradient.norm(2, -1, keepdim=True) + 1e-6)
            normal_marched = segment_coo(src=(weights.unsqueeze(-1) * normal), index=ray_id,
                                         out=torch.zeros([N, 3], device=rgb.device), reduce='sum')
        else:
            normal_marched = None
        if render_kwargs.get('render_depth', False):
            with torch.no_grad():
                depth = segment_coo(src=(weights * step_id * dist).unsqueeze(-1), index=ray_id,
                                    out=torch.zeros([N, 1], device=rgb.device), reduce='sum').squeeze(-1)
            disp = 1 / depth
        else:
            depth, disp = None, 0
        ret_dict.update({
            'alphainv_cum': alphainv_last, 'weights': weights, 'rgb_marched': k_rgb_marched, 'rgb_marched0': rgb_marched,
            'normal_marched': normal_marched, 'raw_alpha': alpha, 'raw_rgb': rgb, 'depth': depth, 'disp': disp,
            'mask': mask, 'mask_outbbox': mask_outbbox, 'gradient': gradient, 'gradient_error': None, 's_val': s_val,
        })
        return ret_dict

    # ------------------------------------------------------------------ mesh / field queries
    @torch.no_grad()
    def query_sdf_field(self, resolution, x_range=None, smooth=True, sigma=0.5, with_gradient=False, chunk=64 ** 3 * 8):
        """The field part of extract_geometry (lib/voxurf_fine.py:894-910 + lib/dvgo_ori.py:679-693): u = -sdf
        (k=3 Gaussian-smoothed when `smooth`) on a resolution^3 lattice over [xyz_min, xyz_max], restricted to the
        X-slab x_range=(x0, x1) so the lattice can be sharded across GPUs.  Marching cubes stays on the host."""
        dev = self.sdf.grid.device
        if self.smooth_sdf:
            sdf_grid = self.smooth_conv(self.sdf.grid)
        elif smooth:
            sdf_grid = SmoothConv(3, sigma)(self.sdf.grid)
        else:
            sdf_grid = self.sdf.grid
        x0, x1 = x_range if x_range is not None else (0, resolution)
        xs = torch.linspace(self._min_host[0], self._max_host[0], resolution, device=dev)[x0:x1]
        ys = torch.linspace(self._min_host[1], self._max_host[1], resolution, device=dev)
        zs = torch.linspace(self._min_host[2], self._max_host[2], resolution, device=dev)
        u = torch.empty(x1 - x0, resolution, resolution, dtype=torch.float32, device=dev)
        g = torch.empty(x1 - x0, resolution, resolution, 3, dtype=torch.float32, device=dev) if with_gradient else None
        planes = max(1, chunk // (resolution * resolution))
        for a in range(0, x1 - x0, planes):
            b = min(a + planes, x1 - x0)
            xx, yy, zz = torch.meshgrid(xs[a:b], ys, zs, indexing='ij')
            pts = torch.stack([xx, yy, zz], -1).reshape(-1, 3)
            if with_gradient:
                s, gr, _ = self.grid_sampler(pts, sdf_grid, sample_ret=True, sample_grad=True)
                g[a:b] = gr.reshape(b - a, resolution, resolution, 3)
            else:
                s = self.grid_sampler(pts, sdf_grid)
            u[a:b] = -s.reshape(b - a, resolution, resolution)
        return (u, g) if with_gradient else u
