"""Drop-in for the one torch_scatter function the reference uses: segment_coo(src, index, out, reduce='sum')
with a sorted index (lib/voxurf_fine.py:753-777, lib/voxurf_coarse.py:575-599).  Deterministic (no atomics)."""
import torch

from ._lib import call


class _SegmentCooSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, index, out):
        src2 = src.contiguous().reshape(src.shape[0], -1)
        call('vx_segment_coo_sum', src2, index, src2.shape[0], src2.shape[1], out)
        ctx.save_for_backward(index)
        ctx.mark_dirty(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (index,) = ctx.saved_tensors
        return grad_out.index_select(0, index), None, None


def segment_coo(src, index, out=None, reduce='sum'):
    if reduce != 'sum':
        raise NotImplementedError("only reduce='sum' is on the Voxurf path")
    if out is None:
        raise ValueError('out= is required (the reference always passes a zeros tensor)')
    if not out.is_cuda:
        out = out.to(src.device)
    return _SegmentCooSum.apply(src, index.contiguous(), out)
