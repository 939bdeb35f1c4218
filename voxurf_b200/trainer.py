"""The loop body of the reference trainer for the surf stages (run.py:552-671), as a reusable step:
forward -> losses (run.py:604-636) -> backward -> total-variation add-grad (run.py:641-655) -> per-voxel Adam
(lib/utils.py:83-199).  bench.py, smoke() and the tests drive this; the reference's run.py can equally drive the
model mirrors itself.

`train_cfg` keys follow configs/dtu_e2e/fine.py / coarse.py (`surf_train`): weight_main, weight_entropy_last,
weight_rgb0, tv_every, tv_from, tv_end, tv_dense_before, weight_tv_density, weight_tv_k0, tv_terms{sdf_tv,
smooth_grad_tv}, ori_tv, lrate_* ...
"""
import torch
import torch.nn.functional as F

from .optim import create_optimizer_or_freeze_model

FINE_TRAIN = dict(   # configs/default_fine_s.py:62-88 (surf_train) overlaid with configs/dtu_e2e/fine.py:21-58
    N_rand=8192, weight_main=1.0, weight_entropy_last=0.001, weight_rgbper=0.0, weight_rgb0=0.5, tv_every=3, tv_from=0,
    tv_end=30000, tv_dense_before=20000, weight_tv_density=0.01, weight_tv_k0=0.0,
    tv_terms=dict(sdf_tv=0.1, grad_tv=0, grad_norm=0, smooth_grad_tv=0.05), ori_tv=False, lrate_decay=20,
    lrate_sdf=5e-3, lrate_k0=1e-1, lrate_rgbnet=3e-3, lrate_k_rgbnet=1e-3, N_iters=20000)

COARSE_TRAIN = dict(  # configs/dtu_e2e/coarse.py:21-58
    N_rand=8192, weight_main=1.0, weight_entropy_last=0.001, weight_rgbper=0.0, weight_rgb0=0.0, tv_every=1, tv_from=0,
    tv_end=20000, tv_dense_before=20000, weight_tv_density=0.001, weight_tv_k0=0.01,
    tv_terms=dict(sdf_tv=0.1, grad_tv=0, smooth_grad_tv=0.05), ori_tv=True, lrate_decay=20, lrate_sdf=0.1, lrate_k0=1e-1,
    lrate_rgbnet=1e-3, N_iters=10000)


class Trainer:
    def __init__(self, model, train_cfg, render_kwargs, zero_grad_in_step=True, grad_sync=None):
        """grad_sync: optional callable(model) run between backward and the TV/Adam step (data-parallel exchange)."""
        self.model, self.cfg, self.render_kwargs = model, dict(train_cfg), dict(render_kwargs)
        self.optimizer = create_optimizer_or_freeze_model(model, self.cfg, global_step=0, zero_grad_in_step=zero_grad_in_step)
        self.zero_grad_in_step = zero_grad_in_step
        self.grad_sync = grad_sync
        self.global_batch = None   # len(rays_o) summed over ranks; run.py:648 divides the TV weight by it

    def is_tv_iter(self, global_step):
        c = self.cfg
        return c['tv_from'] < global_step < c['tv_end'] and global_step % c['tv_every'] == 0

    def loss(self, ret, target, global_step):
        """run.py:604-636"""
        c, model = self.cfg, self.model
        loss = c['weight_main'] * F.mse_loss(ret['rgb_marched'], target)
        if c['weight_entropy_last'] > 0:
            pout = ret['alphainv_cum'][..., -1].clamp(1e-6, 1 - 1e-6)   # last ray only (run.py:608 on a 1-D tensor)
            loss = loss + c['weight_entropy_last'] * (-(pout * torch.log(pout) + (1 - pout) * torch.log(1 - pout)).mean())
        if self.is_tv_iter(global_step) and c['weight_tv_density'] > 0:
            tv = c['tv_terms']
            if tv['smooth_grad_tv'] > 0:
                loss = loss + c['weight_tv_density'] * model.density_total_variation(sdf_tv=0, smooth_grad_tv=tv['smooth_grad_tv'])
            if c.get('ori_tv', False):
                loss = loss + c['weight_tv_density'] * model.density_total_variation(sdf_tv=tv['sdf_tv'], smooth_grad_tv=0)
                if c.get('weight_tv_k0', 0) > 0:
                    loss = loss + c['weight_tv_k0'] * model.k0_total_variation()
        if c.get('weight_rgb0', 0.) > 0 and 'rgb_marched0' in ret:
            loss = loss + F.mse_loss(ret['rgb_marched0'], target) * c['weight_rgb0']
        return loss

    def step(self, rays_o, rays_d, viewdirs, target, global_step):
        c, model = self.cfg, self.model
        tv_iter = self.is_tv_iter(global_step)
        ret = model(rays_o, rays_d, viewdirs, global_step=global_step, materialize_gradient=tv_iter, **self.render_kwargs)
        if not self.zero_grad_in_step:
            self.optimizer.zero_grad(set_to_none=True)
        loss = self.loss(ret, target, global_step)
        loss.backward()
        if self.grad_sync is not None:
            self.grad_sync(model)
        n_batch = self.global_batch or len(rays_o)
        if tv_iter and not c.get('ori_tv', False):   # run.py:641-655
            dense = global_step < c['tv_dense_before']
            if c['weight_tv_density'] > 0 and c['tv_terms']['sdf_tv'] > 0:
                model.sdf_total_variation_add_grad(c['weight_tv_density'] * c['tv_terms']['sdf_tv'] / n_batch, dense)
            if c.get('weight_tv_k0', 0) > 0:
                model.k0_total_variation_add_grad(c['weight_tv_k0'] / n_batch, dense)
        self.optimizer.step()
        return loss.detach(), ret
