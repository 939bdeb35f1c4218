"""Autograd-aware Python faces of the fused operators (the host-side mirror of the ATen calls the
reference makes between its native ops: F.grid_sample, nn.Conv3d, slicing, elementwise chains).

Every function here launches hand-written CUDA through the C ABI (voxurf_b200._lib.call); backward
passes are explicit kernels as well.  Nothing falls back to PyTorch math.
"""
import torch

from ._lib import call
from . import render_utils_cuda


def _is_channels_last(grid):
    """(1,C,X,Y,Z) tensor stored as (X,Y,Z,C) (torch.channels_last_3d) and C > 1."""
    return grid.shape[1] > 1 and not grid.is_contiguous() and grid.is_contiguous(memory_format=torch.channels_last_3d)


def _grid_args(grid):
    if grid.dim() != 5 or grid.shape[0] != 1:
        raise RuntimeError('grid must be (1,C,X,Y,Z)')
    cl = _is_channels_last(grid)
    if not cl and not grid.is_contiguous():
        raise RuntimeError('grid must be contiguous (channel-major or channels_last_3d)')
    return grid.shape[2], grid.shape[3], grid.shape[4], grid.shape[1], int(cl)


def _dense_storage(grid):
    return grid.permute(0, 2, 3, 4, 1) if _is_channels_last(grid) else grid


class _GridGather(torch.autograd.Function):
    """DenseGrid.forward / grid_sampler: lib/grid.py:47-58, lib/voxurf_coarse.py:435-452."""

    @staticmethod
    def forward(ctx, grid, xyz, xyz_min, xyz_max):
        X, Y, Z, C, cl = _grid_args(grid)
        xyz = xyz.reshape(-1, 3).contiguous()
        P = xyz.shape[0]
        out = torch.empty(P, C, dtype=torch.float32, device=xyz.device)
        call('vx_grid_gather', _dense_storage(grid), X, Y, Z, C, cl, xyz_min, xyz_max, xyz, None, None, None, None, 0.0,
             None, P, out)
        ctx.save_for_backward(xyz)
        ctx.geom = (X, Y, Z, C, cl, xyz_min, xyz_max, grid.shape)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        (xyz,) = ctx.saved_tensors
        X, Y, Z, C, cl, xyz_min, xyz_max, shape = ctx.geom
        gg = torch.empty(shape, dtype=torch.float32, device=xyz.device,
                         memory_format=torch.channels_last_3d if cl else torch.contiguous_format).zero_()
        call('vx_grid_gather_backward', X, Y, Z, C, cl, xyz_min, xyz_max, xyz, None, None, None, None, 0.0, None,
             xyz.shape[0], grad_out.contiguous(), _dense_storage(gg), None)
        return gg, None, None, None


def grid_gather(grid, xyz, xyz_min, xyz_max):
    """grid (1,C,X,Y,Z), xyz (...,3) -> (P,C).  xyz_min/xyz_max: host float lists."""
    return _GridGather.apply(grid, xyz, xyz_min, xyz_max)


class _SdfTaps(torch.autograd.Function):
    """Voxurf.grid_sampler(sample_grad=True) and Voxurf.sample_sdfs: lib/voxurf_fine.py:502-577."""

    @staticmethod
    def forward(ctx, grid, xyz, xyz_min, xyz_max, displace, voxel_size, use_grad_norm, xyz_order, want_sdf):
        X, Y, Z, C, cl = _grid_args(grid)
        if C != 1:
            raise RuntimeError('sdf_taps needs a single-channel grid')
        xyz = xyz.reshape(-1, 3).contiguous()
        P, L, dev = xyz.shape[0], len(displace), xyz.device
        sdf = torch.empty(P, dtype=torch.float32, device=dev) if want_sdf else None
        feat = torch.empty(P, 6 * L, dtype=torch.float32, device=dev)
        grad = torch.empty(P, 3 * L, dtype=torch.float32, device=dev)
        call('vx_sdf_taps', grid, X, Y, Z, xyz_min, xyz_max, xyz, None, None, None, None, 0.0, None, P, displace, L,
             voxel_size, int(use_grad_norm), int(xyz_order), sdf, feat, grad)
        ctx.save_for_backward(grid, xyz)
        ctx.cfg = (X, Y, Z, xyz_min, xyz_max, tuple(displace), voxel_size, int(use_grad_norm), int(xyz_order), want_sdf)
        if want_sdf:
            return sdf, feat, grad
        return feat, grad

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *grads):
        grid, xyz = ctx.saved_tensors
        X, Y, Z, xyz_min, xyz_max, displace, voxel_size, ugn, xyz_order, want_sdf = ctx.cfg
        if want_sdf:
            g_sdf, g_feat, g_grad = grads
        else:
            g_sdf, (g_feat, g_grad) = None, grads
        gg = torch.zeros_like(grid)
        c = lambda t: None if t is None else t.contiguous()
        call('vx_sdf_taps_backward', grid, X, Y, Z, xyz_min, xyz_max, xyz, None, None, None, None, 0.0, None,
             xyz.shape[0], list(displace), len(displace), voxel_size, ugn, xyz_order, c(g_sdf), c(g_feat), c(g_grad), gg)
        return gg, None, None, None, None, None, None, None, None


def sdf_taps(grid, xyz, xyz_min, xyz_max, displace, voxel_size, use_grad_norm=False, xyz_order=False, want_sdf=False):
    return _SdfTaps.apply(grid, xyz, xyz_min, xyz_max, [float(d) for d in displace], float(voxel_size),
                          bool(use_grad_norm), bool(xyz_order), bool(want_sdf))


class _NeusAlpha(torch.autograd.Function):
    """neus_alpha_from_sdf_scatter: lib/voxurf_fine.py:475-500."""

    @staticmethod
    def forward(ctx, viewdirs, ray_id, sdf, gradient, dist, inv_s):
        viewdirs, sdf, gradient = viewdirs.contiguous(), sdf.contiguous(), gradient.contiguous()
        n = sdf.shape[0]
        alpha = torch.empty_like(sdf)
        r32, r64 = (ray_id, None) if ray_id.dtype == torch.int32 else (None, ray_id)
        call('vx_neus_alpha', viewdirs, r32, r64, sdf, gradient, dist, inv_s, None, n, alpha, None)
        ctx.save_for_backward(viewdirs, ray_id, sdf, gradient)
        ctx.cfg = (dist, inv_s)
        return alpha

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_alpha):
        viewdirs, ray_id, sdf, gradient = ctx.saved_tensors
        dist, inv_s = ctx.cfg
        g_sdf, g_grad = torch.empty_like(sdf), torch.empty_like(gradient)
        r32, r64 = (ray_id, None) if ray_id.dtype == torch.int32 else (None, ray_id)
        call('vx_neus_alpha_backward', viewdirs, r32, r64, sdf, gradient, dist, inv_s, None, sdf.shape[0],
             grad_alpha.contiguous(), 0, g_sdf, g_grad, None)
        return None, None, g_sdf, g_grad, None, None


def neus_alpha(viewdirs, ray_id, sdf, gradient, dist, inv_s):
    return _NeusAlpha.apply(viewdirs, ray_id.contiguous(), sdf, gradient, float(dist), float(inv_s))


class Alphas2Weights(torch.autograd.Function):
    """lib/voxurf_fine.py:981-997 (== lib/voxurf_coarse.py:717-733)."""

    @staticmethod
    def forward(ctx, alpha, ray_id, N):
        weights, T, alphainv_last, i_start, i_end = render_utils_cuda.alpha2weight(alpha.contiguous(), ray_id, N)
        if alpha.requires_grad:
            ctx.save_for_backward(alpha, weights, T, alphainv_last, i_start, i_end)
            ctx.n_rays = N
        return weights, alphainv_last

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_weights, grad_last):
        alpha, weights, T, alphainv_last, i_start, i_end = ctx.saved_tensors
        grad = render_utils_cuda.alpha2weight_backward(alpha, weights, T, alphainv_last, i_start, i_end, ctx.n_rays,
                                                       grad_weights, grad_last)
        return grad, None, None


class _FdGradient(torch.autograd.Function):
    """neus_sdf_gradient(mode='interpolate'): lib/voxurf_fine.py:440-450."""

    @staticmethod
    def forward(ctx, sdf, voxel_size):
        X, Y, Z, C, cl = _grid_args(sdf)
        out = torch.empty(1, 3, X, Y, Z, dtype=torch.float32, device=sdf.device)
        call('vx_fd_gradient', sdf, X, Y, Z, voxel_size, out)
        ctx.cfg = (X, Y, Z, voxel_size)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        X, Y, Z, voxel_size = ctx.cfg
        ds = torch.zeros(1, 1, X, Y, Z, dtype=torch.float32, device=g.device)
        call('vx_fd_gradient_backward', g.contiguous(), X, Y, Z, voxel_size, ds)
        return ds, None


def fd_gradient(sdf, voxel_size):
    return _FdGradient.apply(sdf.contiguous(), float(voxel_size))


class _Conv3dReplicate(torch.autograd.Function):
    """Conv3d(1,1,k,padding=k//2,padding_mode='replicate'), frozen weights: lib/voxurf_fine.py:246-258."""

    @staticmethod
    def forward(ctx, x, weight, ksize, weight1d):
        B, X, Y, Z = x.shape[0] * x.shape[1], x.shape[2], x.shape[3], x.shape[4]
        out = torch.empty_like(x)
        if weight1d is not None:     # separable kernel: three 1-D passes
            scratch = torch.empty(2 * x.numel(), dtype=torch.float32, device=x.device)
            call('vx_conv3d_replicate_separable', x, B, X, Y, Z, weight1d, ksize, 0, 0, scratch, out)
        else:
            call('vx_conv3d_replicate', x, B, X, Y, Z, weight, ksize, out)
        ctx.cfg = (B, X, Y, Z, weight, ksize, weight1d)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        B, X, Y, Z, weight, ksize, weight1d = ctx.cfg
        din = torch.empty_like(g)
        if weight1d is not None:
            scratch = torch.empty(2 * g.numel(), dtype=torch.float32, device=g.device)
            call('vx_conv3d_replicate_separable', g.contiguous(), B, X, Y, Z, weight1d, ksize, 1, 0, scratch, din)
        else:
            call('vx_conv3d_replicate_backward', g.contiguous(), B, X, Y, Z, weight, ksize, 0, din)
        return din, None, None, None


def conv3d_replicate(x, weight_host, ksize, weight1d=None):
    """x (B,1,X,Y,Z) or (1,B,X,Y,Z): every leading slice is convolved independently; weight_host: flat list k^3;
    weight1d: the k 1-D weights when the kernel is separable (w1 (x) w1 (x) w1)."""
    return _Conv3dReplicate.apply(x.contiguous(), weight_host, int(ksize), weight1d)


class _SmoothGradTV(torch.autograd.Function):
    """density_total_variation(smooth_grad_tv): lib/voxurf_fine.py:417-420; input is the FD gradient grid."""

    @staticmethod
    def forward(ctx, gradient, mask, weight3, w_over_3n):
        X, Y, Z = gradient.shape[2:]
        dG = torch.empty_like(gradient)
        scratch = torch.empty(int(call('vx_smooth_grad_tv_scratch_floats')), dtype=torch.float32, device=gradient.device)
        loss = torch.empty(1, dtype=torch.float32, device=gradient.device)
        call('vx_smooth_grad_tv', gradient, mask, X, Y, Z, weight3, w_over_3n, dG, scratch, loss)
        ctx.save_for_backward(dG)
        return loss[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (dG,) = ctx.saved_tensors
        return dG * g, None, None, None


def smooth_grad_tv(gradient, nonempty_mask, weight3_host, smooth_grad_tv_weight, n_mask):
    return _SmoothGradTV.apply(gradient.contiguous(), nonempty_mask.contiguous(), weight3_host,
                               float(smooth_grad_tv_weight) / (3.0 * float(n_mask)))


class _TotalVariationL1(torch.autograd.Function):
    """total_variation(v, mask): lib/voxurf_fine.py:956-969."""

    @staticmethod
    def forward(ctx, v, mask, inv_cnt):
        C, X, Y, Z = v.shape[1:]
        grad = torch.empty_like(v)
        scratch = torch.empty(3 * int(call('vx_smooth_grad_tv_scratch_floats')), dtype=torch.float32, device=v.device)
        loss = torch.empty(1, dtype=torch.float32, device=v.device)
        call('vx_total_variation_l1', v, mask, C, X, Y, Z, inv_cnt, grad, scratch, loss)
        ctx.save_for_backward(grad)
        return loss[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


def tv_pair_counts(mask, shape):
    """Number of neighbour pairs per axis with both voxels inside `mask` ((X,Y,Z) bool or None). One-off, host side."""
    X, Y, Z = shape
    if mask is None:
        return [(X - 1) * Y * Z, X * (Y - 1) * Z, X * Y * (Z - 1)]
    return [int((mask[:-1] & mask[1:]).sum()), int((mask[:, :-1] & mask[:, 1:]).sum()),
            int((mask[:, :, :-1] & mask[:, :, 1:]).sum())]


def total_variation_l1(v, mask, pair_counts):
    """v (1,C,X,Y,Z) channel-major; mask (X,Y,Z) bool or None; pair_counts from tv_pair_counts (per channel)."""
    inv = [1.0 / (3.0 * max(c, 1) * v.shape[1]) for c in pair_counts]
    return _TotalVariationL1.apply(v.contiguous(), mask, inv)
