"""GPU: the sync-free fused step (voxurf_b200.fused.FusedFineStep) against the CPU oracle and against the drop-in
autograd path -- losses, grid / MLP gradients, TV iterations, Adam updates over several steps, and render()."""
import numpy as np
import pytest
import torch

from oracle import kernels as K
from oracle import voxurf_ref as R
from voxurf_b200 import synthetic as S
from tests.helpers import T, oracle_fine_model, product_fine_model

pytestmark = pytest.mark.gpu
DEV = 'cuda'
RK = dict(near=0.3, far=6.0, bg=0.0, stepsize=0.5)


def close(a, b, rtol=1e-5, atol=1e-6, msg=''):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=msg)


def close_mostly(a, b, atol, hard, msg='', max_frac=1e-4):
    """Parameters after several Adam steps from two runs whose fp32 atomics land in different orders: Adam's first
    steps are sign-like (|dp| ~ lr whatever |g| is), so a voxel whose tiny gradient differs in the last bits can move by
    a sizeable fraction of lr in one run and not in the other.  Require all but a 1e-4 fraction of the elements within
    `atol` and every element within `hard`."""
    d = (a.detach().float().cpu() - b.detach().float().cpu()).abs()
    frac = float((d > atol).float().mean())
    assert frac <= max_frac and float(d.max()) <= hard, (msg, frac, float(d.max()))


def grad_close(a, b, msg='', flips=1e-4, tol=1e-4):
    """|a - b| <= 1e-4 |b| + 1e-4 max|b|, except for a `flips` fraction of the elements (and never beyond 5e-2 max|b|): a
    ReLU gate of a pre-activation within ~1e-6 of zero may flip between two fp32 GEMM formulations,
    which changes the gradient of that one MLP row -- its k0 / sdf taps, a rank-one term of the weight gradients."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    scale = max(float(b.abs().max()), 1e-30)
    d = (a - b).abs()
    bad = d > (tol * b.abs() + tol * scale)
    assert int(bad.sum()) <= max(2, flips * bad.numel()) and float(d.max()) <= 5e-2 * scale, (msg, int(bad.sum()), float(d.max()), scale)


@pytest.mark.parametrize('G,C,cl,n_rays,bg', [(48, 6, True, 1024, 0.0), (40, 12, True, 640, 1.0), (32, 6, False, 512, 0.0)])
def test_fused_forward_backward_matches_oracle(G, C, cl, n_rays, bg):
    from voxurf_b200.fused import FusedFineStep
    from voxurf_b200.trainer import FINE_TRAIN
    sc = S.make_fine_scene(G, C, 64, seed=G + C)
    m = product_fine_model(sc, k0_channels_last=cl)
    om = oracle_fine_model(sc)
    ro, rd, vd = (T(x) for x in S.make_rays(n_rays, seed=G))
    target = T(S.make_target(vd.numpy()))
    rk = dict(RK, bg=bg)
    fs = FusedFineStep(m, n_rays, FINE_TRAIN, rk, row_capacity=4096)
    step = 15003   # TV iteration
    fs.calibrate(ro.to(DEV), rd.to(DEV), vd.to(DEV), global_step=step)
    loss = fs.forward_backward(ro.to(DEV), rd.to(DEV), vd.to(DEV), target.to(DEV), step)
    oret = R.fine_forward(om, ro, rd, vd, step, near=0.3, stepsize=0.5, bg=bg)
    oloss = R.fine_loss(oret, target)
    M0, M2, M4 = fs.counts()
    assert M0 == oret['mask_outbbox'].shape[0] and M2 == int((~oret['mask_outbbox']).sum()) and M4 == oret['weights'].shape[0]
    close(loss, oloss, 1e-5, 1e-7)
    close(fs.rgb_marched, oret['rgb_marched'], 1e-5, 3e-6); close(fs.rgb_marched0, oret['rgb_marched0'], 1e-5, 3e-6)
    close(fs.alphainv_last, oret['alphainv_cum'], 1e-5, 1e-6)
    oloss.backward()
    grad_close(m.sdf.grid.grad, om['sdf'].grad, 'grad_sdf'); grad_close(m.k0.grid.grad, om['k0'].grad, 'grad_k0')
    for mlp, ol in ((fs.mlp1, om['rgbnet']), (fs.mlp2, om['k_rgbnet'])):
        for l, (W, b) in zip(mlp.linears, ol):
            grad_close(l.weight.grad, W.grad, 'W', tol=4e-4); grad_close(l.bias.grad, b.grad, 'b', tol=4e-4)
    if cl:   # every voxel that received a k0 gradient has its `touched` bit set (the sparse-aware Adam relies on it)
        bits = np.unpackbits(fs.k0_touched.cpu().numpy().view(np.uint8), bitorder='little')[:G ** 3].astype(bool)
        nz = (m.k0.grid.grad[0].permute(1, 2, 3, 0).reshape(G ** 3, C) != 0).any(1).cpu().numpy()
        assert (bits | ~nz).all() and bits.sum() <= 8 * M4 and not fs.k0_live.any()
    # regularisers on top (smooth-grad TV through the FD gradient, TV add-grad), then Adam
    fs.regularise(step)
    om['sdf'].grad = None
    (oloss.detach() * 0 + 0.01 * R.smooth_grad_tv(R.sdf_gradient_grid(om['sdf'], om['voxel_size']), om['nonempty_mask'], 0.05)).backward()
    # (first backward's grads were dropped above; compare the regulariser's own contribution)
    g_reg = om['sdf'].grad.clone().contiguous()
    w = 0.01 * 0.1 / n_rays * G / 128
    base = torch.zeros_like(g_reg)
    K.total_variation_add_grad(om['sdf'].detach().contiguous(), base, w, w, w, True)
    g_reg = g_reg + base
    oret2 = R.fine_forward(om, ro, rd, vd, step, near=0.3, stepsize=0.5, bg=bg)
    om['sdf'].grad = None
    R.fine_loss(oret2, target).backward()
    grad_close(m.sdf.grid.grad, om['sdf'].grad + g_reg, 'grad_sdf + regularisers')


def test_fused_steps_match_dropin_autograd_path():
    """Three iterations (one of them a TV iteration): identical parameters from both execution paths."""
    from voxurf_b200.fused import FusedFineStep
    from voxurf_b200.trainer import FINE_TRAIN, Trainer
    sc = S.make_fine_scene(40, 6, 64, seed=21)
    ma, mb = product_fine_model(sc, k0_channels_last=True), product_fine_model(sc, k0_channels_last=True)
    n_rays = 768
    tr = Trainer(ma, FINE_TRAIN, RK, zero_grad_in_step=False)
    fs = FusedFineStep(mb, n_rays, FINE_TRAIN, RK, row_capacity=8192)
    for it, step in enumerate([15001, 15002, 15003]):
        ro, rd, vd = (T(x).to(DEV) for x in S.make_rays(n_rays, seed=100 + it))
        target = T(S.make_target(vd.cpu().numpy(), seed=it)).to(DEV)
        la, _ = tr.step(ro, rd, vd, target, step)
        lb = fs.step(ro, rd, vd, target, step)
        close(lb, la, 1e-5, 1e-7, f'loss step {step}')
        fs.counts()
        # Adam's first steps are sign-like (|dp| ~ lr): compare with an absolute floor of a fraction of lr
        close_mostly(mb.sdf.grid, ma.sdf.grid, 2e-2 * 5e-3, 2 * 5e-3, f'sdf step {step}')
        close_mostly(mb.k0.grid, ma.k0.grid, 2e-2 * 1e-1, 2 * 1e-1, f'k0 step {step}')
        for la_, lb_ in zip([x for x in ma.rgbnet.modules() if isinstance(x, torch.nn.Linear)], fs.mlp1.linears):
            close_mostly(lb_.weight, la_.weight, 5e-2 * 3e-3, 2 * 3e-3, 'rgbnet W', max_frac=1e-3)
    assert (mb.sdf.grid.grad == 0).all() and (mb.k0.grid.grad == 0).all()   # zeroed inside the Adam pass


def test_fused_render_matches_dropin_forward():
    from voxurf_b200.fused import FusedFineStep
    sc = S.make_fine_scene(48, 12, 64, seed=33)
    m = product_fine_model(sc, k0_channels_last=True)
    n_rays = 900
    ro, rd, vd = (T(x).to(DEV) for x in S.make_rays(n_rays, seed=5))
    rk = dict(RK, bg=1.0)
    with torch.no_grad():
        ref = m(ro, rd, vd, render_grad=True, render_depth=True, **rk)
    fs = FusedFineStep(m, n_rays, None, rk, row_capacity=4096)
    fs.calibrate(ro, rd, vd)
    out = fs.render(ro, rd, vd)
    for k in ['rgb_marched', 'rgb_marched0', 'normal_marched', 'depth', 'alphainv_cum']:
        close(out[k], ref[k], 1e-5, 3e-6, k)
    # the same chunk as one CUDA-graph replay (what bench.py --workload render times), on two different ray sets
    fg = FusedFineStep(m, n_rays, None, rk, row_capacity=fs.cap4, use_graph=True)
    for seed in (5, 6, 7):
        ro2, rd2, vd2 = (T(x).to(DEV) for x in S.make_rays(n_rays, seed=seed))
        a = {k: v.clone() for k, v in fs.render(ro2, rd2, vd2).items() if torch.is_tensor(v)}
        b = fg.render_chunk(ro2, rd2, vd2)
        for k in ['rgb_marched', 'rgb_marched0', 'normal_marched', 'depth', 'alphainv_cum']:
            assert torch.equal(a[k], b[k]), k
    assert fg.launches_replayed > 0


@pytest.mark.parametrize('defer', [False, True])
def test_fused_step_cuda_graph_replay_matches_eager(defer):
    """use_graph=True (one CUDA-graph launch per step, step-dependent scalars read from device memory) follows the eager
    step: same losses and parameters over 9 iterations that include TV iterations, replays of both graph variants, a
    changing NeuS s_val, Adam bias corrections and a decaying learning rate."""
    from voxurf_b200.fused import FusedFineStep
    from voxurf_b200.trainer import FINE_TRAIN
    sc = S.make_fine_scene(40, 12, 64, seed=23)
    ma, mb = product_fine_model(sc, k0_channels_last=True), product_fine_model(sc, k0_channels_last=True)
    n_rays = 640
    fa = FusedFineStep(ma, n_rays, FINE_TRAIN, RK, row_capacity=8192)
    fb = FusedFineStep(mb, n_rays, FINE_TRAIN, RK, row_capacity=8192, use_graph=True, defer_optimizer=defer)
    assert fb.use_graph
    for it in range(11 if defer else 9):
        step = 15001 + it
        ro, rd, vd = (T(x).to(DEV) for x in S.make_rays(n_rays, seed=300 + it))
        target = T(S.make_target(vd.cpu().numpy(), seed=it)).to(DEV)
        la = fa.step(ro, rd, vd, target, step).clone()
        lb = fb.step(ro, rd, vd, target, step).clone()
        fa.apply_lr_decay(); fb.apply_lr_decay()
        close(lb, la, 2e-5, 1e-7, f'loss step {step}')
        assert fa.counts() == fb.counts()
    fb.sync_params()      # (defer: the last step's optimizer phase is still pending)
    assert len(fb._graphs) == (3 if defer else 2) and fa.adam_steps == fb.adam_steps == (11 if defer else 9)
    lr = FINE_TRAIN
    close_mostly(mb.sdf.grid, ma.sdf.grid, 2e-2 * lr['lrate_sdf'], 2 * lr['lrate_sdf'], 'sdf')
    close_mostly(mb.k0.grid, ma.k0.grid, 2e-2 * lr['lrate_k0'], 2 * lr['lrate_k0'], 'k0')
    for la_, lb_ in zip(fa.mlp1.linears, fb.mlp1.linears):
        close_mostly(lb_.weight, la_.weight, 5e-2 * lr['lrate_rgbnet'], 2 * lr['lrate_rgbnet'], 'rgbnet W', max_frac=1e-3)
    fb.sync_s_val()
    close(mb.s_val, ma.s_val, 1e-6, 0)


# ------------------------------------------------------------------------------------------------
# fused coarse-stage step (voxurf_b200.fused_coarse.FusedCoarseStep; lib/voxurf_coarse.py:513-619, BASELINE config 2)
# ------------------------------------------------------------------------------------------------
def _coarse_oracle_loss(om, oret, target, c):
    import torch.nn.functional as F
    loss = c['weight_main'] * F.mse_loss(oret['rgb_marched'], target)
    pout = oret['alphainv_cum'][..., -1].clamp(1e-6, 1 - 1e-6)
    loss = loss + c['weight_entropy_last'] * (-(pout * torch.log(pout) + (1 - pout) * torch.log(1 - pout)).mean())
    tv = c['tv_terms']
    reg = c['weight_tv_density'] * R.smooth_grad_tv(oret['_full_gradient'], om['nonempty_mask'], tv['smooth_grad_tv'])
    reg = reg + c['weight_tv_density'] * (R.total_variation_coarse(om['sdf'], om['nonempty_mask']) / 2 / om['voxel_size'] * tv['sdf_tv'])
    reg = reg + c['weight_tv_k0'] * R.total_variation_coarse(om['k0'], om['nonempty_mask'].repeat(1, om['k0'].shape[1], 1, 1, 1))
    return loss, reg


@pytest.mark.parametrize('G,n_rays,bg', [(48, 1024, 0.0), (32, 300, 1.0)])
def test_fused_coarse_step_matches_oracle(G, n_rays, bg):
    from tests.helpers import oracle_coarse_model, product_coarse_model
    from voxurf_b200.fused_coarse import FusedCoarseStep
    from voxurf_b200.trainer import COARSE_TRAIN
    sc = S.make_coarse_scene(G, 12, 128, seed=G)
    m = product_coarse_model(sc)
    om = oracle_coarse_model(sc)
    ro, rd, vd = (T(x) for x in S.make_rays(n_rays, seed=G + 1))
    target = T(S.make_target(vd.numpy()))
    rk = dict(RK, bg=bg)
    fs = FusedCoarseStep(m, n_rays, COARSE_TRAIN, rk, row_capacity=8192)
    step = 1200
    fs.calibrate(ro.to(DEV), rd.to(DEV), vd.to(DEV), global_step=step)
    loss = fs.forward_backward(ro.to(DEV), rd.to(DEV), vd.to(DEV), target.to(DEV), step).clone()
    oret = R.coarse_forward(om, ro, rd, vd, step, near=0.3, stepsize=0.5, bg=bg)
    oloss, oreg = _coarse_oracle_loss(om, oret, target, COARSE_TRAIN)
    M0, M2, M4 = fs.counts()
    assert M0 == oret['mask_outbbox'].shape[0] and M2 == int((~oret['mask_outbbox']).sum()) and M4 == oret['weights'].shape[0]
    close(loss, oloss, 1e-5, 1e-7)
    close(fs.rgb_marched, oret['rgb_marched'], 1e-5, 3e-6); close(fs.alphainv_last, oret['alphainv_cum'], 1e-5, 1e-6)
    fs.regularise(step)
    close(fs.loss, oloss + oreg, 1e-5, 1e-7)
    (oloss + oreg).backward()
    grad_close(m.sdf.grid.grad, om['sdf'].grad, 'grad_sdf'); grad_close(m.k0.grid.grad, om['k0'].grad, 'grad_k0')
    for l, (W, b) in zip(fs.mlp.linears, om['rgbnet']):
        grad_close(l.weight.grad, W.grad, 'W', tol=4e-4); grad_close(l.bias.grad, b.grad, 'b', tol=4e-4)


def test_fused_coarse_graph_replay_matches_dropin_path():
    """Six iterations: the CUDA-graph replayed fused coarse step against the drop-in autograd model + Trainer."""
    from tests.helpers import product_coarse_model
    from voxurf_b200.fused_coarse import FusedCoarseStep
    from voxurf_b200.trainer import COARSE_TRAIN, Trainer
    sc = S.make_coarse_scene(40, 12, 128, seed=7)
    ma, mb = product_coarse_model(sc), product_coarse_model(sc)
    n_rays = 512
    tr = Trainer(ma, COARSE_TRAIN, RK, zero_grad_in_step=False)
    fs = FusedCoarseStep(mb, n_rays, COARSE_TRAIN, RK, row_capacity=16384, use_graph=True)
    for it in range(6):
        step = 1201 + it
        ro, rd, vd = (T(x).to(DEV) for x in S.make_rays(n_rays, seed=400 + it))
        target = T(S.make_target(vd.cpu().numpy(), seed=it)).to(DEV)
        la, _ = tr.step(ro, rd, vd, target, step)
        lb = fs.step(ro, rd, vd, target, step).clone()
        close(lb, la, 2e-5, 1e-7, f'loss step {step}')
        fs.counts()
    assert fs._graph is not None and fs.launches_replayed > 0
    lr = COARSE_TRAIN
    close_mostly(mb.sdf.grid, ma.sdf.grid, 2e-2 * lr['lrate_sdf'], 6 * lr['lrate_sdf'], 'sdf', max_frac=1e-3)
    close_mostly(mb.k0.grid, ma.k0.grid, 2e-2 * lr['lrate_k0'], 6 * lr['lrate_k0'], 'k0', max_frac=1e-3)


def test_deterministic_mode_gives_bit_identical_gradients_and_trajectories():
    """deterministic=True (SURVEY.md 8e "Determinism", north_star "deterministic reductions instead of naive atomics"): every
    backward scatter -- k0 rows, sdf taps, split-K weight gradients -- accumulates in 64-bit fixed point.  The same step run
    twice gives BIT-IDENTICAL sdf / k0 / MLP gradients; eight training steps run twice give bit-identical parameters; and
    the gradients agree with the fp32-atomics path to its own tolerance."""
    from voxurf_b200.fused import FusedFineStep
    from voxurf_b200.trainer import FINE_TRAIN
    sc = S.make_fine_scene(48, 12, 64, seed=91)
    n_rays = 2048
    ro, rd, vd = (T(x).to(DEV) for x in S.make_rays(n_rays, seed=17))
    target = T(S.make_target(vd.cpu().numpy())).to(DEV)

    def grads(det):
        m = product_fine_model(sc, k0_channels_last=True)
        fs = FusedFineStep(m, n_rays, FINE_TRAIN, RK, row_capacity=32768, deterministic=det)
        fs.forward_backward(ro, rd, vd, target, 15003)
        fs.counts()
        return [m.sdf.grid.grad.clone(), m.k0.grid.grad.clone(), fs.mlp_grads.clone()]

    a, b, c = grads(True), grads(True), grads(False)
    for x, y, name in zip(a, b, ('sdf', 'k0', 'mlp')):
        assert torch.equal(x, y), f'{name} gradient differs between two deterministic runs'
        assert float(x.abs().max()) > 0
    for x, y, name in zip(a, c, ('sdf', 'k0', 'mlp')):
        grad_close(x, y, name + ' vs fp32 atomics', tol=2e-5 if name != 'mlp' else 1e-4)

    def train(det):
        m = product_fine_model(sc, k0_channels_last=True)
        fs = FusedFineStep(m, n_rays, FINE_TRAIN, RK, row_capacity=32768, deterministic=det, use_graph=True)
        for it in range(8):
            o, d, v = (T(x).to(DEV) for x in S.make_rays(n_rays, seed=500 + it))
            fs.step(o, d, v, T(S.make_target(v.cpu().numpy(), seed=it)).to(DEV), 15001 + it)
        fs.counts()
        return [m.sdf.grid.detach().clone(), m.k0.grid.detach().clone(), fs.mlp1.flat.detach().clone(), fs.mlp2.flat.detach().clone()]

    pa, pb = train(True), train(True)
    for x, y, name in zip(pa, pb, ('sdf', 'k0', 'rgbnet', 'k_rgbnet')):
        assert torch.equal(x, y), f'{name} parameters differ between two deterministic training runs'
