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
from ._base import SmoothConv, VoxurfBase, _binomial_weights, _gaussian_weights, _mlp  # noqa: F401
from .rays import (batch_indices_generator, get_rays, get_rays_np, get_rays_of_a_view, get_training_rays,  # noqa: F401
                   get_training_rays_flatten, get_training_rays_in_maskcache_sampling, ndc_rays)


class Voxurf(VoxurfBase):
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
        self._init_common(xyz_min, xyz_max, num_voxels, num_voxels_base, alpha_init, s_ratio, s_start, s_learn,
                          step_start, fast_color_thres, nearest)
        self.smooth_scale = smooth_scale
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
        self._init_mask_cache(mask_cache_path, mask_cache_thres, mask_cache_state)
        self.grad_mode = grad_mode

    # ------------------------------------------------------------------ construction helpers
    def get_kwargs(self):
        """lib/voxurf_fine.py:326-342"""
        return {'xyz_min': self.xyz_min.cpu().numpy(), 'xyz_max': self.xyz_max.cpu().numpy(),
                'num_voxels': self.num_voxels, 'num_voxels_base': self.num_voxels_base, 'alpha_init': self.alpha_init,
                'nearest': self.nearest, 'mask_cache_path': self.mask_cache_path,
                'mask_cache_thres': self.mask_cache_thres, 'fast_color_thres': self.fast_color_thres,
                'grad_feat': self.grad_feat, 'sdf_feat': self.sdf_feat, 'k_grad_feat': self.k_grad_feat,
                'k_sdf_feat': self.k_sdf_feat, **self.rgbnet_kwargs}

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

    # ------------------------------------------------------------------ regularisers
    def density_total_variation(self, sdf_tv=0, smooth_grad_tv=0, grad_tv=0, smooth_sdf_tv=0):
        """lib/voxurf_fine.py:412-421 (the smooth_grad_tv branch is the one the shipped configs use)."""
        tv = 0
        if sdf_tv > 0:
            raise NotImplementedError('autograd-form sdf_tv is a coarse-stage (ori_tv) option; see voxurf_coarse')
        if smooth_grad_tv > 0:
            tv = tv + ops.smooth_grad_tv(self.gradient, self.nonempty_mask[0, 0], self._tv_smooth_w, smooth_grad_tv,
                                         self._n_nonempty)
        return tv

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
        s_val, inv_s = self._update_s_val(global_step)
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
    def query_sdf_field(self, resolution, x_range=None, smooth=True, sigma=0.5, with_gradient=False, sdf_grid=None):
        """The field part of extract_geometry (lib/voxurf_fine.py:894-910 + lib/dvgo_ori.py:679-693): u = -sdf
        (k=3 Gaussian-smoothed when `smooth`) on a resolution^3 lattice over [xyz_min, xyz_max], restricted to the
        X-slab x_range=(x0, x1) so the lattice can be sharded across GPUs (grids are replicated, the smoothing halo is
        local).  with_gradient: also the (.., 3) gradient of the same grid at the lattice points (the 6-tap
        grid_sampler gradient, lib/voxurf_fine.py:502-534).  One fused launch (vx_sdf_lattice): no points are
        materialised.  sdf_grid: pass the already smoothed grid when querying several slabs."""
        dev = self.sdf.grid.device
        if sdf_grid is None:
            sdf_grid = self.mesh_query_grid(smooth, sigma)
        x0, x1 = x_range if x_range is not None else (0, resolution)
        # torch.linspace on the host in float32, like extract_fields builds its axes; three tiny uploads
        ax = [torch.linspace(self._min_host[i], self._max_host[i], resolution) for i in range(3)]
        xs, ys, zs = ax[0][x0:x1].contiguous().to(dev), ax[1].to(dev), ax[2].to(dev)
        u = torch.empty(x1 - x0, resolution, resolution, dtype=torch.float32, device=dev)
        g = torch.empty(x1 - x0, resolution, resolution, 3, dtype=torch.float32, device=dev) if with_gradient else None
        X, Y, Z = (int(w) for w in sdf_grid.shape[2:])
        call('vx_sdf_lattice', sdf_grid, X, Y, Z, self._min_host, self._max_host, xs, ys, zs, x1 - x0, resolution, resolution,
             self._voxel_size_host, 1, u, g)
        return (u, g) if with_gradient else u

    @torch.no_grad()
    def mesh_color_forward(self, ray_pts, **kwargs):
        """lib/voxurf_fine.py:804-892: colours of mesh vertices (viewdirs = -normal).  Runs on the fused step's kernels
        (FusedFineStep.mesh_colors: vx_sdf_taps, vx_fused_row_features, the tcgen05 MLP chains)."""
        fs = getattr(self, '_fused', None)
        if fs is None:
            from .fused import FusedFineStep
            fs = FusedFineStep(self, 8, None, dict(near=0.0, stepsize=0.5), row_capacity=65536)
        return fs.mesh_colors(ray_pts)

    def extract_geometry(self, bound_min=None, bound_max=None, resolution=128, threshold=0.0, smooth=True, sigma=0.5, **kwargs):
        """lib/voxurf_fine.py:894-910: (vertices, triangles) of the iso-surface -sdf = threshold; field query and marching
        cubes both on the GPU (voxurf_b200/marching.py).  Returns numpy arrays like the reference."""
        from . import marching
        if resolution is None:
            resolution = int(self.world_size[0])
        v, t = marching.extract_geometry(self, resolution, threshold, smooth, sigma)
        return v.cpu().numpy(), t.cpu().numpy()

    @torch.no_grad()
    def mesh_query_grid(self, smooth=True, sigma=0.5):
        """The grid extract_geometry queries: the per-iteration smoothed grid if the model has one, else the k=3 Gaussian
        smoothed grid (init_smooth_conv_test_k3), else the raw grid (lib/voxurf_fine.py:894-903)."""
        if self.smooth_sdf:
            return self.smooth_conv(self.sdf.grid)
        if smooth:
            return SmoothConv(3, sigma).to(self.sdf.grid.device)(self.sdf.grid)
        return self.sdf.grid
