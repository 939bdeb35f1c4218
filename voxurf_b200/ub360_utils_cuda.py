"""Drop-in for the reference's `ub360_utils_cuda` module (lib/cuda/ub360_utils.cpp:17-25), the one native op of the
unbounded "womask" models (lib/voxurf_womask_fine.py:867, lib/voxurf_womask_coarse.py:645)."""
import torch

from ._lib import call


def cumdist_thres(dist, thres):
    """dist (n_rays, n_pts) float32 CUDA, contiguous -> bool mask of the same shape"""
    if not dist.is_cuda:
        raise RuntimeError('dist must be a CUDA tensor')
    if not dist.is_contiguous():
        raise RuntimeError('dist must be contiguous')
    mask = torch.zeros(dist.shape, dtype=torch.bool, device=dist.device)
    call('vx_cumdist_thres', dist, float(thres), dist.shape[0], dist.shape[1], mask)
    return mask
