"""Drop-in for the reference's `total_variation_cuda` module (lib/cuda/total_variation.cpp:29-32)."""
from ._lib import call


def _shape5(param):
    if param.dim() != 5:
        raise RuntimeError('param must be (1,C,X,Y,Z)')
    return param.shape[2], param.shape[3], param.shape[4]


def total_variation_add_grad(param, grad, wx, wy, wz, dense_mode):
    i, j, k = _shape5(param)
    call('vx_total_variation_add_grad', param, grad, None, wx, wy, wz, int(bool(dense_mode)), i, j, k, param.numel())


def total_variation_add_grad_new(param, grad, mask, wx, wy, wz, dense_mode):
    i, j, k = _shape5(param)
    call('vx_total_variation_add_grad', param, grad, mask, wx, wy, wz, int(bool(dense_mode)), i, j, k, param.numel())
