"""Host-side mirror of the reference's optimizer, lib/utils.py:54-229 (`Adam` with optional per-voxel
learning rate and `create_optimizer_or_freeze_model`), on the fused one-pass kernel vx_adam_step.

The reference's `adam()` makes ~8 elementwise passes per parameter tensor (lib/utils.py:185-199); the
kernel reads p, g, m, v once and writes p, m, v once, and can zero the gradient in the same pass
(`zero_grad_in_step=True`) so a persistent `.grad` buffer replaces the per-iteration zero-allocation.
"""
import math

import torch
import torch.nn as nn

from ._lib import call


def _storage(t):
    """Dense view of a tensor's storage (handles channels_last_3d grids)."""
    if t.is_contiguous():
        return t
    if t.dim() == 5 and t.is_contiguous(memory_format=torch.channels_last_3d):
        return t.permute(0, 2, 3, 4, 1)
    raise RuntimeError('Adam: parameter must be dense (contiguous or channels_last_3d)')


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False,
                 zero_grad_in_step=False):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError('weight_decay / amsgrad are never enabled by the reference (lib/utils.py:229)')
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad)
        self.per_lr = None
        self.zero_grad_in_step = zero_grad_in_step
        self.timed_param, self.timings = None, []   # bench.py: CUDA events around one parameter's kernel
        super().__init__(params, defaults)

    def set_pervoxel_lr(self, count):
        """lib/utils.py:79-81"""
        assert self.param_groups[0]['params'][0].shape == count.shape
        self.per_lr = (count.float() / count.max()).contiguous()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    continue
                state = self.state[p]
                if len(state) == 0:
                    state['step'] = 0
                    state['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state['step'] += 1
                step = state['step']
                bias_correction1 = 1 - beta1 ** step           # lib/utils.py:176-177 (python doubles)
                bias_correction2 = 1 - beta2 ** step
                step_size = group['lr'] / bias_correction1     # :192
                per_lr = self.per_lr if (self.per_lr is not None and p.shape == self.per_lr.shape) else None
                g = p.grad
                if not (g.is_contiguous() or g.is_contiguous(memory_format=torch.channels_last_3d)) or \
                        g.stride() != p.stride():
                    g = g.contiguous(memory_format=torch.preserve_format) if g.stride() == p.stride() else \
                        torch.empty_like(p).copy_(g)
                timed = p is self.timed_param
                if timed:
                    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                    ev[0].record()
                call('vx_adam_step', _storage(p.data), _storage(g), _storage(state['exp_avg']),
                     _storage(state['exp_avg_sq']), per_lr, p.numel(), beta1, beta2, 1 - beta1, 1 - beta2, step_size,
                     math.sqrt(bias_correction2), group['eps'], 0, int(self.zero_grad_in_step), None, None, 1, None)
                if self.zero_grad_in_step and g is not p.grad:
                    p.grad.zero_()   # the kernel zeroed the dense copy, not the strided gradient autograd accumulates into
                if timed:
                    ev[1].record()
                    self.timings.append(ev)
        return loss


def create_optimizer_or_freeze_model(model, cfg_train, global_step, zero_grad_in_step=False):
    """lib/utils.py:202-229: one param group per `lrate_<name>` key of cfg_train (dict or attribute bag)."""
    get = (lambda k: cfg_train[k]) if isinstance(cfg_train, dict) else (lambda k: getattr(cfg_train, k))
    decay_steps = get('lrate_decay') * 1000
    decay_factor = 0.1 ** (global_step / decay_steps)
    param_group = []
    for k in list(cfg_train.keys()):
        if not k.startswith('lrate_') or k == 'lrate_decay':
            continue
        name = k[len('lrate_'):]
        if not hasattr(model, name):
            continue
        param = getattr(model, name)
        if param is None:
            continue
        lr = get(k) * decay_factor
        if lr > 0:
            if isinstance(param, nn.Module):
                param = param.parameters()
            param_group.append({'params': param, 'lr': lr, 'name': name})
        else:
            param.requires_grad = False
    return Adam(param_group, betas=(0.9, 0.99), zero_grad_in_step=zero_grad_in_step)
