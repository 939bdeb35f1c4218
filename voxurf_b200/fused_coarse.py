"""Sync-free fused coarse-stage step: Voxurf.forward + losses + backward + autograd-form regularisers + Adam of the 96^3
coarse stage (lib/voxurf_coarse.py:513-619, run.py:600-659 with configs/dtu_e2e/coarse.py: ori_tv=True, tv_every=1,
per-iteration 5^3 smoothing) on persistent device buffers, no host syncs, one CUDA-graph replay per step.

The drop-in autograd mirror of the same model (voxurf_coarse.Voxurf.forward) takes ~6 ms per step on cuBLAS + ATen;
this is the B200-first execution of the same arithmetic.  Sequence (kernel -> reference lines):
  vx_ray_setup, vx_march_flags/emit     bbox + MaskCache -> (ray_id, step_id)             voxurf_coarse.py:454-486,524-530
  vx_conv3d_replicate_separable         per-iteration Gaussian smoothing of the sdf grid  :531 (smooth_conv)
  vx_fd_gradient                        gradient GRID of the RAW sdf grid (hazard 11)      :534 (neus_sdf_gradient)
  vx_grid_gather x2                     sdf from the smoothed grid, 3-channel gradient     :533,535
  vx_neus_alpha                         NeuS alpha                                        :538
  vx_alpha2weight_seg x2                weights, weight > thres, then AGAIN on survivors   :540-550 (hazard 8)
  vx_scan_i32, vx_fused_emit_rows       row list
  vx_coarse_row_features                k0 gather, PEs, normal -> X                        :552-569
  vx_mlp_chain_batch                    rgbnet 57 -> 128 -> 128 -> 3 (tcgen05, TF32x3; last layer on CUDA cores) :570
  vx_coarse_composite_loss              sigmoid, segment sums, (1 - sum w) bg, clamp, mse, entropy  :573-583, run.py:604-610
  vx_mlp_chain_batch (dX), vx_mlp_dw_batch
  vx_alpha2weight_seg_backward, vx_neus_alpha_backward, vx_coarse_row_backward
  vx_grid_gather_backward x2, conv adjoint, vx_fd_gradient_backward
  vx_smooth_grad_tv, vx_total_variation_l1 x2     autograd-form regularisers, run.py:612-625, voxurf_coarse.py:300-320,702-715
  vx_adam_step x3                        sdf, k0, rgbnet                                   lib/utils.py:154-199
"""
import math

import numpy as np
import torch

from ._lib import call
from .mlp import FlatMLP, prepare_chains, run_chain_jobs, run_dw_batch
from .optim import _storage


class FusedCoarseStep:
    def __init__(self, model, n_rays, train_cfg=None, render_kwargs=None, row_capacity=262144, use_graph=False):
        m = model
        if m.k0_dim not in (6, 12):
            raise NotImplementedError('fused coarse step: k0 channels must be 6 or 12')
        if m.k0.channels_last:
            raise NotImplementedError('fused coarse step: the k0 regulariser (voxurf_coarse.py:311-320) needs the channel-major k0 layout')
        if m.mask_cache is None:
            raise NotImplementedError('fused coarse step needs a mask cache (the shipped coarse configs have one)')
        self.m, self.N = m, int(n_rays)
        self.cfg = dict(train_cfg) if train_cfg is not None else None
        self.rk = dict(render_kwargs or {})
        dev = m.sdf.grid.device
        self.dev = dev
        self.X, self.Y, self.Z = (int(w) for w in m.world_size)
        self.C = m.k0_dim
        self.P, self.V = m.posfreq.numel(), m.viewfreq.numel()
        self.D = self.C + 3 + 6 * self.P + 3 + 6 * self.V + 3
        self.ld = (self.D + 15) // 16 * 16
        self.stepdist = float(np.float32(float(self.rk.get('stepsize', 0.5)) * m._voxel_size_host))
        self.dist = self.stepdist
        N = self.N
        cap2 = N * m._max_steps(self.stepdist)
        self.cap2 = cap2
        f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        i32 = lambda *s: torch.empty(*s, dtype=torch.int32, device=dev)
        u8 = lambda *s: torch.empty(*s, dtype=torch.uint8, device=dev)
        self.t_min, self.t_max = f32(N), f32(N)
        self.n_steps = torch.empty(N, dtype=torch.int64, device=dev)
        self.start, self.dirs = f32(N, 3), f32(N, 3)
        self.offsets = torch.empty(N + 1, dtype=torch.int64, device=dev)
        words = N * (m._max_steps(self.stepdist) // 32 + 2) + 1
        self.bits_in, self.bits_keep = i32(words), i32(words)
        self.keep_count, self.keep_off = i32(N), i32(N + 1)
        self.w_count, self.off4 = i32(N), i32(N + 1)
        self.w_count1 = i32(N)
        self.alphainv_last, self.alphainv_last1 = f32(N), f32(N)
        self.i_end, self.i_end1, self.d_last, self.loss_ray = i32(N), i32(N), f32(N), f32(N)
        self.rgb_marched = f32(N, 3)
        self.loss = f32(1)
        self.ray_id, self.step_id = i32(cap2), i32(cap2)
        self.sdf_s, self.grad_s, self.alpha = f32(cap2), f32(cap2, 3), f32(cap2)
        self.w_keep1, self.w_keep = u8(cap2), u8(cap2)
        self.weight1, self.T1, self.weight, self.T = f32(cap2), f32(cap2), f32(cap2), f32(cap2)
        self.d_w, self.d_alpha, self.d_sdf_s, self.d_grad_s = torch.zeros(cap2, device=dev), f32(cap2), f32(cap2), f32(cap2, 3)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        # grids
        self.smoothed, self.d_smoothed = torch.empty_like(m.sdf.grid), torch.zeros_like(m.sdf.grid)
        self.conv_scratch = torch.empty(2 * m.sdf.grid.numel(), dtype=torch.float32, device=dev)
        self.G = torch.empty(1, 3, self.X, self.Y, self.Z, dtype=torch.float32, device=dev)
        self.dG = torch.zeros_like(self.G)
        self.sdf_grad = torch.zeros_like(m.sdf.grid)
        self.k0_grad = torch.zeros_like(m.k0.grid)
        m.sdf.grid.grad, m.k0.grid.grad = self.sdf_grad, self.k0_grad
        self.tv_grad_sdf, self.tv_grad_k0 = torch.empty_like(m.sdf.grid), torch.empty_like(m.k0.grid)
        n_scr = int(call('vx_smooth_grad_tv_scratch_floats'))
        self.tv_scratch = torch.empty(3 * n_scr, dtype=torch.float32, device=dev)
        self.tv_loss = torch.zeros(3, dtype=torch.float32, device=dev)      # smooth-grad TV, sdf TV, k0 TV
        # MLP
        self.cap4 = int(row_capacity)
        self.mlp = FlatMLP(m.rgbnet, self.ld, self.D, True)
        self._alloc_rows(self.cap4)
        self.adam_state, self.adam_steps = {}, 0
        self.use_graph = bool(use_graph)
        self._graph, self._eager_done, self._dev_consts = None, False, None
        self._graph_launches, self.launches_replayed = 0, 0
        self.consts = torch.zeros(16, dtype=torch.float32, device=dev)
        if self.use_graph:
            self.in_o, self.in_d, self.in_v, self.in_t = (torch.zeros(n_rays, 3, dtype=torch.float32, device=dev) for _ in range(4))
        if self.cfg is not None:
            c = self.cfg
            self.groups = [('sdf', m.sdf.grid, c['lrate_sdf']), ('k0', m.k0.grid, c['lrate_k0']), ('rgbnet', self.mlp.flat, c['lrate_rgbnet'])]
            self.lr = {name: lr for name, _, lr in self.groups}

    def _alloc_rows(self, cap):
        dev = self.dev
        self.cap4 = cap
        self.idx4 = torch.empty(cap, dtype=torch.int32, device=dev)
        self.Xr = torch.empty(cap, self.ld, dtype=torch.float32, device=dev)
        self.dXr = torch.empty_like(self.Xr)
        self.logit = torch.empty(cap, 3, dtype=torch.float32, device=dev)
        self.d_logit = torch.zeros_like(self.logit)
        self.mlp.alloc(cap)

    def _geom(self):
        m = self.m
        return (self.X, self.Y, self.Z, m._min_host, m._max_host)

    def _pts(self):
        return (self.ray_id, self.step_id, self.start, self.dirs, self.stepdist)

    # ------------------------------------------------------------------ forward
    def _forward(self, rays_o, rays_d, viewdirs, global_step, train):
        m, N = self.m, self.N
        assert rays_o.shape[0] == N
        rays_o, rays_d, viewdirs = rays_o.contiguous(), rays_d.contiguous(), viewdirs.contiguous()
        if self._dev_consts is None:
            s_val, inv_s = m._update_s_val(global_step)
            inv_s_dev = None
        else:
            s_val, inv_s, inv_s_dev = 0, 0.0, self._dev_consts[0:1]
        self.inv_s, self._inv_s_dev = inv_s, inv_s_dev
        X, Y, Z, mn, mx = self._geom()
        call('vx_ray_setup', rays_o, rays_d, m.xyz_min, m.xyz_max, self.rk['near'], 1e9, self.stepdist, N, self.t_min, self.t_max,
             self.n_steps, self.start, self.dirs, self.offsets)
        call('vx_march_flags_cells', self.start, self.dirs, m.xyz_min, m.xyz_max, self.offsets, N, self.stepdist, *m.mask_cache.march_args_cells(),
             self.bits_in, self.bits_keep, self.keep_count, self.keep_off)
        call('vx_march_emit', self.offsets, N, self.bits_keep, self.keep_off, self.cap2, self.ray_id, self.step_id, None)
        n2 = self.keep_off[N:]
        if m.smooth_sdf:
            call('vx_conv3d_replicate_separable', m.sdf.grid, 1, X, Y, Z, m.smooth_conv.weight1d_host, m.smooth_conv.ksize, 0, 0,
                 self.conv_scratch, self.smoothed)
            sdf_grid = self.smoothed
        else:
            sdf_grid = m.sdf.grid
        call('vx_fd_gradient', m.sdf.grid, X, Y, Z, m._voxel_size_host, self.G)
        m.gradient = self.G
        call('vx_grid_gather', sdf_grid, X, Y, Z, 1, 0, mn, mx, None, *self._pts(), n2, 0, self.sdf_s)
        call('vx_grid_gather', self.G, X, Y, Z, 3, 0, mn, mx, None, *self._pts(), n2, 0, self.grad_s)
        call('vx_neus_alpha', viewdirs, self.ray_id, None, self.sdf_s, self.grad_s, self.dist, inv_s, n2, 0, self.alpha, inv_s_dev)
        thres = float(m.fast_color_thres)
        if thres > 0:      # :540-548: weights of all samples, weight > thres ...
            call('vx_alpha2weight_seg', self.alpha, None, self.keep_off, N, thres, self.weight1, self.T1, self.alphainv_last1,
                 self.i_end1, self.w_keep1, self.w_count1)
            keep = self.w_keep1
        else:
            keep = None
        # ... and alpha2weight AGAIN on the survivors (:550): their weights, T and alphainv_last are what is composited
        call('vx_alpha2weight_seg', self.alpha, keep, self.keep_off, N, -1.0, self.weight, self.T, self.alphainv_last, self.i_end,
             self.w_keep, self.w_count)
        self._keep = keep
        call('vx_scan_i32', self.w_count, N, self.off4)
        call('vx_fused_emit_rows', self.w_keep, self.keep_off, self.off4, N, self.cap4, self.idx4, self.overflow)
        n4 = self.off4[N:]
        call('vx_coarse_row_features', _storage(m.k0.grid), X, Y, Z, self.C, 0, mn, mx, *self._pts(), self.idx4, n4, self.cap4,
             viewdirs, self.grad_s, self.P, self.V, self.ld, self.Xr)
        prepare_chains(self.mlp.chains(train))
        self.mlp._n = n4
        run_chain_jobs([self.mlp.forward_job(self.Xr, self.logit, train)], n4, self.cap4, None)
        return s_val, n2, n4

    @torch.no_grad()
    def render(self, rays_o, rays_d, viewdirs):
        s_val, n2, n4 = self._forward(rays_o, rays_d, viewdirs, None, False)
        call('vx_coarse_composite_loss', self.logit, 3, self.idx4, self.off4, self.cap4, self.weight, self.alphainv_last, None, self.N,
             0.0, 0.0, 0.0, float(self.rk.get('bg', 0.0)), 0, self.rgb_marched, None, None, None, None)
        return {'rgb_marched': self.rgb_marched, 'alphainv_cum': self.alphainv_last, 's_val': s_val}

    @torch.no_grad()
    def forward_backward(self, rays_o, rays_d, viewdirs, target, global_step):
        """forward + data losses + backward into the persistent gradient buffers (regularisers: regularise())"""
        m, N, c = self.m, self.N, self.cfg or {}
        s_val, n2, n4 = self._forward(rays_o, rays_d, viewdirs, global_step, True)
        X, Y, Z, mn, mx = self._geom()
        call('vx_coarse_composite_loss', self.logit, 3, self.idx4, self.off4, self.cap4, self.weight, self.alphainv_last,
             target.contiguous(), N, c.get('weight_main', 1.0), c.get('weight_entropy_last', 0.0), 1.0, float(self.rk.get('bg', 0.0)), 1,
             self.rgb_marched, self.d_logit, self.d_w, self.d_last, self.loss_ray)
        call('vx_sum_f32', self.loss_ray, N, self.loss)
        # M2-level backward first (it writes d_sdf_s / d_grad_s of every sample), then the rows add the normal's adjoint
        call('vx_alpha2weight_seg_backward', self.alpha, self.weight, self.T, self._keep, self.alphainv_last, self.keep_off, self.i_end,
             N, self.d_w, self.d_last, self.d_alpha)
        call('vx_neus_alpha_backward', viewdirs.contiguous(), self.ray_id, None, self.sdf_s, self.grad_s, self.dist, self.inv_s, n2, 0,
             self.d_alpha, 0, self.d_sdf_s, self.d_grad_s, self._inv_s_dev)
        run_chain_jobs([self.mlp.backward_job(self.d_logit, self.dXr)], n4, self.cap4, None)
        run_dw_batch([self.mlp])
        call('vx_coarse_row_backward', X, Y, Z, self.C, 0, mn, mx, *self._pts(), self.idx4, n4, self.cap4, self.grad_s, self.P, self.V,
             self.ld, self.dXr, self.d_grad_s, _storage(self.k0_grad))
        grad_target = self.d_smoothed if m.smooth_sdf else self.sdf_grad
        call('vx_grid_gather_backward', X, Y, Z, 1, 0, mn, mx, None, *self._pts(), n2, 0, self.d_sdf_s, grad_target, None)
        call('vx_grid_gather_backward', X, Y, Z, 3, 0, mn, mx, None, *self._pts(), n2, 0, self.d_grad_s, self.dG, None)
        return self.loss

    def is_tv_iter(self, global_step):
        c = self.cfg
        return c['tv_from'] < global_step < c['tv_end'] and global_step % c['tv_every'] == 0

    @torch.no_grad()
    def regularise(self, global_step):
        """run.py:612-625 with ori_tv=True: smooth-gradient TV, L1 TV of the sdf grid, L1 TV of the k0 grid
        (lib/voxurf_coarse.py:300-320, 702-715), in gradient form; then the grid-level adjoints (FD gradient, smoothing)."""
        c, m = self.cfg, self.m
        X, Y, Z = self.X, self.Y, self.Z
        tv_iter = self.is_tv_iter(global_step) and c['weight_tv_density'] > 0
        self.tv_loss.zero_()
        if tv_iter:
            assert c.get('ori_tv', False), 'fused coarse step: the add-grad TV form belongs to the fine stage'
            tv = c['tv_terms']
            mask3 = m.nonempty_mask[0, 0]
            n_mask = m._n_nonempty
            if tv['smooth_grad_tv'] > 0:   # d/dG lands in dG (the sampled-gradient scatter already accumulated there)
                w = c['weight_tv_density'] * tv['smooth_grad_tv'] / (3.0 * n_mask)
                dG_tv = self.tv_grad_k0.view(-1)[:self.dG.numel()].view_as(self.dG)      # scratch
                call('vx_smooth_grad_tv', self.G, mask3, X, Y, Z, m._tv_smooth_w, w, dG_tv, self.tv_scratch, self.tv_loss[0:1])
                call('vx_axpy', dG_tv, 1.0, self.dG.numel(), self.dG)
            if tv['sdf_tv'] > 0:
                inv = [1.0 / (3.0 * n_mask)] * 3
                call('vx_total_variation_l1', m.sdf.grid, mask3, 1, X, Y, Z, inv, self.tv_grad_sdf, self.tv_scratch, self.tv_loss[1:2])
                s = c['weight_tv_density'] * tv['sdf_tv'] / 2 / m._voxel_size_host
                call('vx_axpy', self.tv_grad_sdf, s, self.sdf_grad.numel(), self.sdf_grad)
                self.tv_loss[1:2].mul_(s)
            if c.get('weight_tv_k0', 0) > 0:
                inv = [1.0 / (3.0 * n_mask * self.C)] * 3
                call('vx_total_variation_l1', m.k0.grid, mask3, self.C, X, Y, Z, inv, self.tv_grad_k0, self.tv_scratch, self.tv_loss[2:3])
                call('vx_axpy', self.tv_grad_k0, c['weight_tv_k0'], self.k0_grad.numel(), self.k0_grad)
                self.tv_loss[2:3].mul_(c['weight_tv_k0'])
        # grid-level adjoints: gradient grid -> raw sdf, smoothed grid -> raw sdf
        call('vx_fd_gradient_backward', self.dG, X, Y, Z, m._voxel_size_host, self.sdf_grad)
        self.dG.zero_()
        if m.smooth_sdf:
            call('vx_conv3d_replicate_separable', self.d_smoothed, 1, X, Y, Z, m.smooth_conv.weight1d_host, m.smooth_conv.ksize, 1, 1,
                 self.conv_scratch, self.sdf_grad)
            self.d_smoothed.zero_()
        call('vx_sum_f32', self.tv_loss, 3, self.consts[15:16])
        self.loss.add_(self.consts[15:16])

    @torch.no_grad()
    def optimizer_step(self):
        if self._dev_consts is None:
            self.adam_steps += 1
        step, beta1, beta2, eps = max(self.adam_steps, 1), 0.9, 0.99, 1e-8
        bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
        for gi, (name, p, _) in enumerate(self.groups):
            st = self.adam_state.get(name)
            if st is None:
                st = (torch.zeros_like(p), torch.zeros_like(p))
                self.adam_state[name] = st
            call('vx_adam_step', _storage(p.data), _storage(p.grad), _storage(st[0]), _storage(st[1]), None, p.numel(), beta1, beta2,
                 1 - beta1, 1 - beta2, self.lr[name] / bc1, math.sqrt(bc2), eps, 0, 1, None, None, 1,
                 None if self._dev_consts is None else self._dev_consts[2 + 2 * gi:4 + 2 * gi])

    def apply_lr_decay(self):
        f = 0.1 ** (1 / (self.cfg['lrate_decay'] * 1000))
        for k in self.lr:
            self.lr[k] *= f

    def _step_body(self, rays_o, rays_d, viewdirs, target, global_step):
        loss = self.forward_backward(rays_o, rays_d, viewdirs, target, global_step)
        self.regularise(global_step if global_step is not None else self._gs_host)
        self.optimizer_step()
        return loss

    def step(self, rays_o, rays_d, viewdirs, target, global_step):
        """One coarse-stage training iteration.  With use_graph the second call captures the step, later calls replay it
        (every iteration is a TV iteration in the shipped config, so there is one variant)."""
        self._gs_host = global_step
        if not self.use_graph or not self._eager_done:
            self._eager_done = True
            return self._step_body(rays_o, rays_d, viewdirs, target, global_step)
        m = self.m
        s_val = 1. / (global_step + m.s_ratio / m.s_start - m.step_start) * m.s_ratio
        m._s_val_host = float(np.float32(s_val))
        self.adam_steps += 1
        step = self.adam_steps
        bc1, bc2 = 1 - 0.9 ** step, 1 - 0.99 ** step
        host = [float(np.float32(1.0) / np.float32(m._s_val_host)), 0.0]
        for name, _, _ in self.groups:
            host += [self.lr[name] / bc1, math.sqrt(bc2)]
        self.consts[:len(host)].copy_(torch.tensor(host, dtype=torch.float32))
        torch._foreach_copy_([self.in_o, self.in_d, self.in_v, self.in_t], [rays_o, rays_d, viewdirs, target])
        if self._graph is None:
            from ._lib import launch_count
            self._tv_variant = self.is_tv_iter(global_step)
            self._dev_consts = self.consts
            g = torch.cuda.CUDAGraph()
            l0 = launch_count()
            try:
                with torch.cuda.graph(g):
                    self._step_body(self.in_o, self.in_d, self.in_v, self.in_t, None)
            finally:
                self._dev_consts = None
            self._graph, self._graph_launches = g, launch_count() - l0
        assert self.is_tv_iter(global_step) == self._tv_variant, 'fused coarse step captured one TV variant (tv_every = 1)'
        self._graph.replay()
        self.launches_replayed += self._graph_launches
        return self.loss

    def counts(self):
        v = torch.stack([self.offsets[self.N], self.keep_off[self.N].long(), self.off4[self.N].long(), self.overflow[0].long()]).cpu()
        if int(v[3]) > 0:
            raise RuntimeError(f'FusedCoarseStep: {int(v[3])} MLP rows exceed row_capacity={self.cap4}; call calibrate() or raise it')
        return int(v[0]), int(v[1]), int(v[2])

    def calibrate(self, rays_o, rays_d, viewdirs, global_step=None, headroom=1.3, multiple=4096):
        self.overflow.zero_()
        with torch.no_grad():
            self._forward(rays_o, rays_d, viewdirs, global_step, False)
        v = torch.stack([self.off4[self.N], self.overflow[0]]).cpu()
        m4 = max(int(v[0]), int(v[1]))
        cap = max(multiple, int(math.ceil(m4 * headroom / multiple)) * multiple)
        self.overflow.zero_()
        if cap != self.cap4:
            self._alloc_rows(cap)
        return cap
