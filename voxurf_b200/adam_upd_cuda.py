"""Drop-in for the reference's `adam_upd_cuda` module (lib/cuda/adam_upd.cpp:79-86)."""
from ._lib import call


def adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    call('vx_adam_upd', param, grad, exp_avg, exp_avg_sq, None, param.numel(), step, beta1, beta2, lr, eps, 0)


def masked_adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    call('vx_adam_upd', param, grad, exp_avg, exp_avg_sq, None, param.numel(), step, beta1, beta2, lr, eps, 1)


def adam_upd_with_perlr(param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps):
    call('vx_adam_upd', param, grad, exp_avg, exp_avg_sq, perlr, param.numel(), step, beta1, beta2, lr, eps, 2)
