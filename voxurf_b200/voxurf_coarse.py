"""Host-side mirror of the reference's coarse-stage Voxurf model, lib/voxurf_coarse.py (96^3 SDF + 12-channel k0 +
one rgbnet), on the B200 operators.  Same constructor keywords, attributes, `forward()` signature and ret_dict
keys as lib/voxurf_coarse.py:25-151,513-619.

Differences from the fine model that matter for parity:
  * the SDF value is sampled from the Gaussian-smoothed grid but the gradient is a 3-channel trilinear gather of
    the central-difference gradient grid of the RAW sdf grid (lib/voxurf_coarse.py:531-535);
  * one weights > thres compaction, then alpha2weight is run AGAIN on the survivors (:540-550);
  * the background is blended with 1 - sum(weights) and the result clamped (:575-583);
  * the regularisers are the autograd forms (ori_tv=True, configs/dtu_e2e/coarse.py:29): L1 total variation of the
    sdf and k0 grids and the smooth-gradient TV (lib/voxurf_coarse.py:300-320).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import grid, ops
from .grid import MaskCache  # noqa: F401
from .ops import Alphas2Weights  # noqa: F401
from .torch_scatter import segment_coo
from ._base import SmoothConv, VoxurfBase, _mlp  # noqa: F401
from .rays import (batch_indices_generator, get_rays, get_rays_np, get_rays_of_a_view, get_training_rays,  # noqa: F401
                   get_training_rays_flatten, get_training_rays_in_maskcache_sampling, ndc_rays)


class Voxurf(VoxurfBase):
    def __init__(self, xyz_min, xyz_max, num_voxels=0, num_voxels_base=0, alpha_init=None, nearest=False,
                 mask_cache_path=None, mask_cache_thres=1e-3, fast_color_thres=0, rgbnet_dim=0, rgbnet_direct=False,
                 rgbnet_full_implicit=False, rgbnet_depth=3, rgbnet_width=128, posbase_pe=5, viewbase_pe=4,
                 geo_rgb_dim=3, grad_mode='interpolate', s_ratio=2000, s_start=0.2, s_learn=False, step_start=0,
                 smooth_ksize=0, smooth_sigma=1, mask_cache_state=None, k0_channels_last=False, **kwargs):
        super().__init__()
        if nearest or s_learn or rgbnet_dim <= 0 or rgbnet_full_implicit or not rgbnet_direct or geo_rgb_dim != 3 \
                or grad_mode != 'interpolate':
            raise NotImplementedError('option outside the configurations the reference ships (configs/*/coarse.py)')
        self._init_common(xyz_min, xyz_max, num_voxels, num_voxels_base, alpha_init, s_ratio, s_start, s_learn,
                          step_start, fast_color_thres, nearest)
        self.init_smooth_conv(smooth_ksize, smooth_sigma)
        self.rgbnet_kwargs = {'rgbnet_dim': rgbnet_dim, 'rgbnet_direct': rgbnet_direct,
                              'rgbnet_full_implicit': rgbnet_full_implicit, 'rgbnet_depth': rgbnet_depth,
                              'rgbnet_width': rgbnet_width, 'posbase_pe': posbase_pe, 'viewbase_pe': viewbase_pe}
        self.rgbnet_full_implicit = rgbnet_full_implicit
        self.k0_dim = rgbnet_dim
        self.k0 = grid.create_grid('DenseGrid', channels=self.k0_dim, world_size=self.world_size, xyz_min=self.xyz_min,
                                   xyz_max=self.xyz_max, channels_last=k0_channels_last)
        self.rgbnet_direct = rgbnet_direct
        self.register_buffer('posfreq', torch.FloatTensor([(2 ** i) for i in range(posbase_pe)]))
        self.register_buffer('viewfreq', torch.FloatTensor([(2 ** i) for i in range(viewbase_pe)]))
        dim0 = (3 + 3 * posbase_pe * 2) + (3 + 3 * viewbase_pe * 2) + self.k0_dim + geo_rgb_dim
        self.geo_rgb_dim = geo_rgb_dim
        self.rgbnet = _mlp(dim0, rgbnet_width, rgbnet_depth)
        nn.init.constant_(self.rgbnet[-1].bias, 0)
        self._init_mask_cache(mask_cache_path, mask_cache_thres, mask_cache_state)
        self.grad_mode = grad_mode
        self._pair_counts = None

    def get_kwargs(self):
        """lib/voxurf_coarse.py:262-275"""
        return {'xyz_min': self.xyz_min.cpu().numpy(), 'xyz_max': self.xyz_max.cpu().numpy(),
                'num_voxels': self.num_voxels, 'num_voxels_base': self.num_voxels_base, 'alpha_init': self.alpha_init,
                'nearest': self.nearest, 'mask_cache_path': self.mask_cache_path,
                'mask_cache_thres': self.mask_cache_thres, 'fast_color_thres': self.fast_color_thres,
                'geo_rgb_dim': self.geo_rgb_dim, **self.rgbnet_kwargs}

    # ------------------------------------------------------------------ regularisers (autograd forms)
    def _mask3(self):
        return None if self.nonempty_mask is None else self.nonempty_mask[0, 0]

    def _counts(self):
        """Normaliser of lib/voxurf_coarse.py:702-715: the three masked L1 sums are divided by 3 * mask.sum() -- the
        number of masked VOXELS (times C through the repeated mask), the same for every axis -- unlike the fine file's
        per-axis pair means (lib/voxurf_fine.py:956-969)."""
        if self._pair_counts is None or self._pair_counts[0] is not self.nonempty_mask:
            n = self._n_nonempty if self.nonempty_mask is not None else int(self.sdf.grid[0, 0].numel())
            self._pair_counts = (self.nonempty_mask, [n, n, n])
        return self._pair_counts[1]

    def density_total_variation(self, sdf_tv=0, smooth_grad_tv=0, sdf_thrd=0.999):
        """lib/voxurf_coarse.py:300-309"""
        tv = 0
        if sdf_tv > 0:
            tv = tv + ops.total_variation_l1(self.sdf.grid, self._mask3(), self._counts()) / 2 / self._voxel_size_host * sdf_tv
        if smooth_grad_tv > 0:
            tv = tv + ops.smooth_grad_tv(self.gradient, self.nonempty_mask[0, 0], self._tv_smooth_w, smooth_grad_tv,
                                         self._n_nonempty)
        return tv

    def k0_total_variation(self, k0_tv=1., k0_grad_tv=0.):
        """lib/voxurf_coarse.py:311-320 (rgbnet is not None -> the raw k0 grid)"""
        if k0_grad_tv > 0:
            raise NotImplementedError
        if self.k0.channels_last:
            raise RuntimeError('k0_total_variation needs the channel-major layout')
        return ops.total_variation_l1(self.k0.grid, self._mask3(), self._counts()) if k0_tv > 0 else 0

    # ------------------------------------------------------------------ samplers
    def grid_sampler(self, xyz, *grids, mode=None, align_corners=True, smooth=False, displace=0.):
        """lib/voxurf_coarse.py:435-452 (displace == 0 on every live call site)"""
        if displace != 0:
            raise NotImplementedError
        g = self.smooth_conv(grids[0]) if smooth else grids[0]
        shape = xyz.shape[:-1]
        return ops.grid_gather(g, xyz, self._min_host, self._max_host).reshape(*shape, g.shape[1]).squeeze()

    def sample_ray_cuda(self, rays_o, rays_d, near, far, stepsize, maskout=True, use_bg=False, **render_kwargs):
        """lib/voxurf_coarse.py:454-486"""
        if use_bg or not maskout:
            raise NotImplementedError
        return self.sample_ray(rays_o, rays_d, near, far, stepsize)

    # ------------------------------------------------------------------ forward
    def forward(self, rays_o, rays_d, viewdirs, global_step=None, **render_kwargs):
        """Volume rendering, lib/voxurf_coarse.py:513-619."""
        ret_dict = {}
        N = len(rays_o)
        viewdirs = viewdirs.contiguous()
        m = self._march(rays_o, rays_d, render_kwargs['near'], render_kwargs['stepsize'])
        ray_pts, step_id, mask_outbbox = m['ray_pts'], m['step_id'].long(), m['mask_outbbox']
        ray_id = m['ray_id'].long()
        mask = None

        sdf_grid = self.smooth_conv(self.sdf.grid) if self.smooth_sdf else self.sdf.grid
        sdf = self.grid_sampler(ray_pts, sdf_grid).reshape(-1)
        self.gradient = self.neus_sdf_gradient(sdf=self.sdf.grid)
        gradient = self.grid_sampler(ray_pts, self.gradient).reshape(-1, 3)
        dist = render_kwargs['stepsize'] * self._voxel_size_host
        s_val, inv_s = self._update_s_val(global_step)
        import numpy as np
        alpha = ops.neus_alpha(viewdirs, ray_id, sdf, gradient, float(np.float32(dist)), inv_s)

        weights, alphainv_last = Alphas2Weights.apply(alpha, ray_id, N)
        if self.fast_color_thres > 0:   # :542-548
            mask = weights > self.fast_color_thres
            ray_pts, ray_id, step_id, alpha, gradient = (t[mask] for t in (ray_pts, ray_id, step_id, alpha, gradient))
        weights, alphainv_last = Alphas2Weights.apply(alpha, ray_id, N)   # :550

        k0 = self.k0(ray_pts)
        rays_xyz = (ray_pts - self.xyz_min) / (self.xyz_max - self.xyz_min)
        xyz_emb = (rays_xyz.unsqueeze(-1) * self.posfreq).flatten(-2)
        xyz_emb = torch.cat([rays_xyz, xyz_emb.sin(), xyz_emb.cos()], -1)
        viewdirs_emb = (viewdirs.unsqueeze(-1) * self.viewfreq).flatten(-2)
        viewdirs_emb = torch.cat([viewdirs, viewdirs_emb.sin(), viewdirs_emb.cos()], -1)
        rgb_feat = torch.cat([k0, xyz_emb, viewdirs_emb.flatten(0, -2)[ray_id]], -1)
        normal = gradient / (gradient.norm(dim=-1, keepdim=True) + 1e-5)
        rgb_feat = torch.cat([rgb_feat, normal], -1)
        rgb_logit = self.rgbnet(rgb_feat)
        rgb = torch.sigmoid(rgb_logit)

        dev = rgb.device
        rgb_marched = segment_coo(src=(weights.unsqueeze(-1) * rgb), index=ray_id, out=torch.zeros([N, 3], device=dev), reduce='sum')
        cum_weights = segment_coo(src=(weights.unsqueeze(-1)), index=ray_id, out=torch.zeros([N, 1], device=dev), reduce='sum')
        rgb_marched = rgb_marched + (1 - cum_weights) * render_kwargs['bg']
        rgb_marched = rgb_marched.clamp(0, 1)
        if gradient is not None and render_kwargs.get('render_grad', False):
            nrm = gradient / (gradient.norm(2, -1, keepdim=True) + 1e-6)
            normal_marched = segment_coo(src=(weights.unsqueeze(-1) * nrm), index=ray_id, out=torch.zeros([N, 3], device=dev), reduce='sum')
        else:
            normal_marched = None
        # `getattr(render_kwargs, 'render_depth', False)` on a dict is always False in the reference (:594)
        ret_dict.update({
            'alphainv_cum': alphainv_last, 'weights': weights, 'rgb_marched': rgb_marched, 'normal_marched': normal_marched,
            'raw_alpha': alpha, 'raw_rgb': rgb, 'depth': None, 'disp': None, 'mask': mask, 'mask_outbbox': mask_outbbox,
            'gradient': gradient, 'gradient_error': None, 's_val': s_val,
        })
        return ret_dict
