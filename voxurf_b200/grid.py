"""Host-side mirror of the reference's lib/grid.py (DenseGrid, MaskGrid, create_grid) and of the
MaskCache module of lib/voxurf_fine.py:917-942, on the B200 operators.

Same constructor arguments, attribute names (`.grid`, `.xyz_min`, `.xyz_max`, `.channels`,
`.world_size`), state_dict keys and method names as the reference, so checkpoints and callers carry
over.  One B200-first extension: `DenseGrid(..., channels_last=True)` keeps a multi-channel grid in
torch.channels_last_3d memory, i.e. (X,Y,Z,C) in HBM with the logical shape still (1,C,X,Y,Z): the 8
corner reads of a C-channel trilinear gather become 8 contiguous vectors instead of 8*C strided scalars.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, render_utils_cuda, total_variation_cuda
from ._lib import call


def _host3(t):
    return [float(v) for v in torch.as_tensor(t).flatten().tolist()]


def create_grid(type, **kwargs):
    if type == 'DenseGrid':
        return DenseGrid(**kwargs)
    raise NotImplementedError(type)  # TensoRFGrid: no reference config selects it


class DenseGrid(nn.Module):
    """lib/grid.py:38-92."""

    def __init__(self, channels, world_size, xyz_min, xyz_max, channels_last=False, **kwargs):
        super().__init__()
        self.channels = channels
        self.world_size = world_size
        self.register_buffer('xyz_min', torch.Tensor(_host3(xyz_min)))
        self.register_buffer('xyz_max', torch.Tensor(_host3(xyz_max)))
        self._min_host, self._max_host = _host3(xyz_min), _host3(xyz_max)
        self.channels_last = bool(channels_last) and channels > 1
        g = torch.zeros([1, channels, *[int(w) for w in world_size]])
        if self.channels_last:
            g = g.contiguous(memory_format=torch.channels_last_3d)
        self.grid = nn.Parameter(g)

    def forward(self, xyz):
        """xyz: global coordinates to query -> (..., C) (squeezed when C == 1), lib/grid.py:47-58"""
        shape = xyz.shape[:-1]
        out = ops.grid_gather(self.grid, xyz, self._min_host, self._max_host).reshape(*shape, self.channels)
        if self.channels == 1:
            out = out.squeeze(-1)
        return out

    def scale_volume_grid(self, new_world_size):
        """lib/grid.py:60-65 (trilinear F.interpolate; progressive growing happens once, off the hot path)."""
        new_world_size = [int(w) for w in new_world_size]
        if self.channels == 0:
            self.grid = nn.Parameter(torch.zeros([1, self.channels, *new_world_size], device=self.grid.device))
        else:
            g = F.interpolate(self.grid.data.contiguous(), size=tuple(new_world_size), mode='trilinear', align_corners=True)
            if self.channels_last:
                g = g.contiguous(memory_format=torch.channels_last_3d)
            self.grid = nn.Parameter(g)
        self.world_size = torch.tensor(new_world_size)

    def total_variation_add_grad(self, wx, wy, wz, dense_mode, mask=None):
        """Add gradients by total variation loss in-place, lib/grid.py:67-78."""
        if self.channels_last:
            raise RuntimeError('total_variation_add_grad needs the channel-major layout')
        if mask is None:
            total_variation_cuda.total_variation_add_grad(self.grid, self.grid.grad, wx, wy, wz, dense_mode)
        else:
            mask = mask.detach()
            if self.grid.size(1) > 1 and mask.size() != self.grid.size():
                mask = mask.repeat(1, self.grid.size(1), 1, 1, 1).contiguous()
            assert mask.size() == self.grid.size()
            total_variation_cuda.total_variation_add_grad_new(self.grid, self.grid.grad, mask.float(), wx, wy, wz, dense_mode)

    def get_dense_grid(self):
        return self.grid

    @torch.no_grad()
    def __isub__(self, val):
        self.grid.data -= val
        return self

    def extra_repr(self):
        return f'channels={self.channels}, world_size={[int(w) for w in self.world_size]}'


class MaskGrid(nn.Module):
    """lib/grid.py:212-245: bool occupancy grid with nearest-voxel lookup."""

    def __init__(self, path=None, mask_cache_thres=None, mask=None, xyz_min=None, xyz_max=None):
        super().__init__()
        if path is not None:
            st = torch.load(path, weights_only=False)
            self.mask_cache_thres = mask_cache_thres
            density = F.max_pool3d(st['model_state_dict']['density.grid'], kernel_size=3, padding=1, stride=1)
            alpha = 1 - torch.exp(-F.softplus(density + st['model_state_dict']['act_shift']) * st['model_kwargs']['voxel_size_ratio'])
            mask = (alpha >= self.mask_cache_thres).squeeze(0).squeeze(0)
            xyz_min = torch.Tensor(st['model_kwargs']['xyz_min'])
            xyz_max = torch.Tensor(st['model_kwargs']['xyz_max'])
        else:
            mask = mask.bool()
            xyz_min = torch.Tensor(_host3(xyz_min))
            xyz_max = torch.Tensor(_host3(xyz_max))
        self.register_buffer('mask', mask.contiguous())
        xyz_len = xyz_max - xyz_min
        self.register_buffer('xyz2ijk_scale', (torch.Tensor(list(mask.shape)) - 1) / xyz_len)
        self.register_buffer('xyz2ijk_shift', -xyz_min * self.xyz2ijk_scale)

    @torch.no_grad()
    def forward(self, xyz):
        shape = xyz.shape[:-1]
        xyz = xyz.reshape(-1, 3).contiguous()
        mask = render_utils_cuda.maskcache_lookup(self.mask, xyz, self.xyz2ijk_scale, self.xyz2ijk_shift)
        return mask.reshape(shape)

    def extra_repr(self):
        return f'mask.shape={list(self.mask.shape)}'


class MaskCache(nn.Module):
    """lib/voxurf_fine.py:917-942: free-space mask from the coarse-stage density checkpoint.
    `path` is a checkpoint with the reference's schema, or pass `state` = the loaded dict directly."""

    def __init__(self, path=None, mask_cache_thres=1e-3, ks=3, state=None):
        super().__init__()
        st = state if state is not None else torch.load(path, weights_only=False)
        self.mask_cache_thres = mask_cache_thres
        kw = st['MaskCache_kwargs']
        self.register_buffer('xyz_min', torch.FloatTensor(_host3(kw['xyz_min'])))
        self.register_buffer('xyz_max', torch.FloatTensor(_host3(kw['xyz_max'])))
        self._min_host, self._max_host = _host3(kw['xyz_min']), _host3(kw['xyz_max'])
        self.register_buffer('density', F.max_pool3d(st['model_state_dict']['density'], kernel_size=ks, padding=ks // 2, stride=1).contiguous())
        self.act_shift = float(kw['act_shift'])
        self.voxel_size_ratio = float(kw['voxel_size_ratio'])
        self.nearest = kw.get('nearest', False)
        if self.nearest:
            raise NotImplementedError('nearest-mode MaskCache is not on the Voxurf path')

    @torch.no_grad()
    def forward(self, xyz):
        shape = xyz.shape[:-1]
        pts = xyz.reshape(-1, 3).contiguous()
        out = torch.empty(pts.shape[0], dtype=torch.bool, device=pts.device)
        d = self.density
        call('vx_mask_cache_query', d, d.shape[2], d.shape[3], d.shape[4], self._min_host, self._max_host,
             self.act_shift, self.voxel_size_ratio, float(self.mask_cache_thres), pts, pts.shape[0], out)
        return out.reshape(shape)

    def march_args_cells(self):
        """march_args() + the per-cell verdict table, for vx_march_flags_cells (built once per device / threshold)."""
        d = self.density
        key = (d.data_ptr(), d._version, float(self.mask_cache_thres))
        if getattr(self, '_cells_key', None) != key:
            self._cells = torch.empty(d.shape[2:], dtype=torch.uint8, device=d.device)
            call('vx_mask_cache_cells', d, d.shape[2], d.shape[3], d.shape[4], self.act_shift, self.voxel_size_ratio,
                 float(self.mask_cache_thres), self._cells)
            self._cells_key = key
        return self.march_args() + (self._cells,)

    def march_args(self):
        """(density, X, Y, Z, min_host, max_host, act_shift, voxel_size_ratio, thres) for vx_march_flags."""
        d = self.density
        return (d, d.shape[2], d.shape[3], d.shape[4], self._min_host, self._max_host, self.act_shift,
                self.voxel_size_ratio, float(self.mask_cache_thres))
