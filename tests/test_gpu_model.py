"""GPU: the drop-in fine model (voxurf_b200.voxurf_fine.Voxurf) end to end -- forward, backward, TV add-grad and
the trainer's Adam -- against (i) the committed golden vectors produced by the reference's own Python and
(ii) the CPU oracle at a larger size."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import kernels as K
from oracle import voxurf_ref as R
from voxurf_b200 import synthetic as S
from tests.helpers import (T, load_golden, oracle_coarse_model, oracle_fine_model, product_coarse_model,
                           product_fine_model)

pytestmark = pytest.mark.gpu
DEV = 'cuda'
RET_KEYS = ['alphainv_cum', 'weights', 'rgb_marched', 'rgb_marched0', 'normal_marched', 'raw_alpha', 'raw_rgb', 'depth',
            'disp', 'mask', 'mask_outbbox', 'gradient']


def close(a, b, rtol=1e-5, atol=1e-6, msg=''):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=msg)


def loss_fn(ret, target):
    loss = F.mse_loss(ret['rgb_marched'], target)
    pout = ret['alphainv_cum'][..., -1].clamp(1e-6, 1 - 1e-6)
    loss = loss + 0.001 * (-(pout * torch.log(pout) + (1 - pout) * torch.log(1 - pout)).mean())
    return loss + 0.5 * F.mse_loss(ret['rgb_marched0'], target)


def grad_close(a, b, rtol=1e-4, rel_floor=1e-4, msg=''):
    """gradient tolerance (north_star: 1e-4 on gradients): |d| <= 1e-4 * |g| + 1e-4 * max|g|.  The floor covers grid
    voxels whose gradient is a cancelling sum of many fp32 atomics (different accumulation order than the oracle)."""
    scale = float(np.abs(b).max()) if not torch.is_tensor(b) else float(b.abs().max())
    close(a, b, rtol, rel_floor * max(scale, 1e-30), msg)


@pytest.mark.parametrize('k0_cl', [False, True])
def test_fine_model_matches_reference_golden(k0_cl):
    g = load_golden('fine_forward.npz')
    sc = S.make_fine_scene(20, 6, 32, seed=3, mask_G=12)
    m = product_fine_model(sc, k0_channels_last=k0_cl)
    close(m.sdf.grid, g['sdf_after_mask'], 0, 0)
    ro, rd, vd = (T(x).to(DEV) for x in S.make_rays(96, seed=777))
    target = T(S.make_target(vd.cpu().numpy())).to(DEV)
    rk = dict(near=0.3, far=6.0, bg=0, stepsize=0.5, render_grad=True, render_depth=True)
    ret = m(ro, rd, vd, global_step=15001, **rk)
    assert ret['s_val'] == float(g['s_val'])
    for k in RET_KEYS:
        if ret[k].dtype == torch.bool:
            assert (ret[k].cpu().numpy() == g[k]).all(), k
        else:
            close(ret[k], g[k], 1e-5, 2e-6, k)
    close(m.gradient, g['full_gradient'], 1e-5, 1e-6)
    loss = loss_fn(ret, target)
    close(loss, g['loss'], 1e-5, 1e-7)
    loss.backward()
    grad_close(m.sdf.grid.grad, g['grad_sdf'], msg='grad_sdf'); grad_close(m.k0.grid.grad, g['grad_k0'], msg='grad_k0')
    for name, net in (('rgbnet', m.rgbnet), ('k_rgbnet', m.k_rgbnet)):
        for i, l in enumerate([x for x in net.modules() if isinstance(x, torch.nn.Linear)]):
            grad_close(l.weight.grad, g[f'grad_{name}_W{i}'], msg=f'{name} W{i}'); grad_close(l.bias.grad, g[f'grad_{name}_b{i}'], msg=f'{name} b{i}')
    # TV add-grad + the trainer's Adam (run.py:641-659)
    if not k0_cl:
        from voxurf_b200.optim import Adam
        m.sdf_total_variation_add_grad(0.01 * 0.1 / 96, True)
        grad_close(m.sdf.grid.grad, g['grad_sdf_after_tv'], msg='tv')
        opt = Adam([{'params': [m.sdf.grid], 'lr': 5e-3}, {'params': [m.k0.grid], 'lr': 1e-1}], betas=(0.9, 0.99))
        opt.step()
        close(m.sdf.grid, g['sdf_after_adam'], 1e-5, 1e-5); close(m.k0.grid, g['k0_after_adam'], 1e-5, 1e-5)
        with torch.no_grad():
            ret_e = m(ro, rd, vd, **rk)
        close(ret_e['rgb_marched'], g['eval_rgb_marched'], 1e-4, 1e-4)
        close(ret_e['normal_marched'], g['eval_normal_marched'], 1e-4, 1e-4)
        close(ret_e['depth'], g['eval_depth'], 1e-4, 1e-4)


@pytest.mark.parametrize('G,n_rays,step,C', [(64, 1024, 15001, 6), (48, 512, 200, 12)])
def test_fine_model_matches_oracle(G, n_rays, step, C):
    sc = S.make_fine_scene(G, C, 64, seed=G)
    m = product_fine_model(sc, k0_channels_last=(C == 12))
    om = oracle_fine_model(sc)
    ro, rd, vd = (T(x) for x in S.make_rays(n_rays, seed=G + 1))
    target = T(S.make_target(vd.numpy()))
    kw = dict(near=0.3, stepsize=0.5, bg=1.0 if C == 12 else 0.0, render_grad=True, render_depth=True)
    oret = R.fine_forward(om, ro, rd, vd, step, **kw)
    ret = m(ro.to(DEV), rd.to(DEV), vd.to(DEV), global_step=step, far=6.0, **kw)
    for k in RET_KEYS:
        if oret[k].dtype == torch.bool:
            assert torch.equal(ret[k].cpu(), oret[k]), k
        else:
            close(ret[k], oret[k], 1e-5, 3e-6, k)
    oloss = R.fine_loss(oret, target)
    loss = loss_fn(ret, target.to(DEV))
    close(loss, oloss, 1e-5, 1e-7)
    # + smooth-grad TV regulariser through the full-grid FD gradient (voxurf_fine.py:412-421, run.py:622-625)
    oloss = oloss + 0.01 * R.smooth_grad_tv(oret['_full_gradient'], om['nonempty_mask'], 0.05)
    loss = loss + 0.01 * m.density_total_variation(sdf_tv=0, smooth_grad_tv=0.05)
    oloss.backward(); loss.backward()
    grad_close(m.sdf.grid.grad, om['sdf'].grad, msg='grad_sdf'); grad_close(m.k0.grid.grad, om['k0'].grad, msg='grad_k0')
    for net, ol in ((m.rgbnet, om['rgbnet']), (m.k_rgbnet, om['k_rgbnet'])):
        for l, (W, b) in zip([x for x in net.modules() if isinstance(x, torch.nn.Linear)], ol):
            grad_close(l.weight.grad, W.grad); grad_close(l.bias.grad, b.grad)


def test_coarse_model_matches_reference_golden():
    g = load_golden('coarse_forward.npz')
    sc = S.make_coarse_scene(16, 12, 32, seed=4, mask_G=12)
    m = product_coarse_model(sc)
    close(m.sdf.grid, g['sdf_after_mask'], 0, 0)
    ro, rd, vd = (T(x).to(DEV) for x in S.make_rays(96, seed=777))
    target = T(S.make_target(vd.cpu().numpy())).to(DEV)
    ret = m(ro, rd, vd, global_step=2000, near=0.3, far=6.0, bg=0, stepsize=0.5, render_grad=True)
    for k in ['alphainv_cum', 'weights', 'rgb_marched', 'normal_marched', 'raw_alpha', 'raw_rgb', 'mask', 'mask_outbbox', 'gradient']:
        if ret[k].dtype == torch.bool:
            assert (ret[k].cpu().numpy() == g[k]).all(), k
        else:
            close(ret[k], g[k], 1e-5, 2e-6, k)
    loss = F.mse_loss(ret['rgb_marched'], target)
    close(loss, g['loss'], 1e-5, 1e-7)
    loss.backward()
    grad_close(m.sdf.grid.grad, g['grad_sdf'], msg='grad_sdf'); grad_close(m.k0.grid.grad, g['grad_k0'], msg='grad_k0')
    for i, l in enumerate([x for x in m.rgbnet.modules() if isinstance(x, torch.nn.Linear)]):
        grad_close(l.weight.grad, g[f'grad_rgbnet_W{i}']); grad_close(l.bias.grad, g[f'grad_rgbnet_b{i}'])


@pytest.mark.parametrize('G,n_rays,cl', [(48, 1024, False), (32, 300, True)])
def test_coarse_model_matches_oracle_with_regularisers(G, n_rays, cl):
    sc = S.make_coarse_scene(G, 12, 64, seed=G)
    m = product_coarse_model(sc, k0_channels_last=cl)
    om = oracle_coarse_model(sc)
    ro, rd, vd = (T(x) for x in S.make_rays(n_rays, seed=G + 1))
    target = T(S.make_target(vd.numpy()))
    kw = dict(near=0.3, stepsize=0.5, bg=1.0, render_grad=True)
    oret = R.coarse_forward(om, ro, rd, vd, 1200, **kw)
    ret = m(ro.to(DEV), rd.to(DEV), vd.to(DEV), global_step=1200, far=6.0, **kw)
    for k in ['alphainv_cum', 'weights', 'rgb_marched', 'normal_marched', 'raw_alpha', 'raw_rgb', 'mask', 'mask_outbbox', 'gradient']:
        if oret[k].dtype == torch.bool:
            assert torch.equal(ret[k].cpu(), oret[k]), k
        else:
            close(ret[k], oret[k], 1e-5, 3e-6, k)
    # run.py:604-628 with ori_tv=True (configs/dtu_e2e/coarse.py:27-45)
    oloss = F.mse_loss(oret['rgb_marched'], target)
    oloss = oloss + 0.001 * R.smooth_grad_tv(oret['_full_gradient'], om['nonempty_mask'], 0.2)
    oloss = oloss + 0.001 * (R.total_variation_coarse(om['sdf'], om['nonempty_mask']) / 2 / om['voxel_size'] * 0.1)
    loss = F.mse_loss(ret['rgb_marched'], target.to(DEV))
    loss = loss + 0.001 * m.density_total_variation(sdf_tv=0, smooth_grad_tv=0.2)
    loss = loss + 0.001 * m.density_total_variation(sdf_tv=0.1, smooth_grad_tv=0)
    if not cl:
        oloss = oloss + 0.01 * R.total_variation_coarse(om['k0'], om['nonempty_mask'].repeat(1, 12, 1, 1, 1))
        loss = loss + 0.01 * m.k0_total_variation()
    close(loss, oloss, 1e-5, 1e-7)
    oloss.backward(); loss.backward()
    grad_close(m.sdf.grid.grad, om['sdf'].grad, msg='grad_sdf'); grad_close(m.k0.grid.grad, om['k0'].grad, msg='grad_k0')
    for l, (W, b) in zip([x for x in m.rgbnet.modules() if isinstance(x, torch.nn.Linear)], om['rgbnet']):
        grad_close(l.weight.grad, W.grad); grad_close(l.bias.grad, b.grad)


def test_query_sdf_field_matches_reference_and_shards_by_slab():
    """Mesh field query (BASELINE config 5 at small size): the smoothed -sdf and its gradient on a lattice, whole and as
    X-slabs, against the vectors of the reference's own extract_fields / grid_sampler (tests/golden/r2_extras.npz) and
    against the oracle at a lattice size that does not divide the grid."""
    g = load_golden('r2_extras.npz')
    sc = S.make_fine_scene(24, 6, 32, seed=5, mask_G=12)
    m = product_fine_model(sc)
    om = oracle_fine_model(sc, requires_grad=False)
    u, gr = m.query_sdf_field(24, smooth=True, sigma=0.5, with_gradient=True)
    close(u, g['field_u'], 1e-5, 2e-6, 'u vs extract_fields'); close(-u, g['field_sdf'], 1e-5, 2e-6)
    close(gr, g['field_grad'], 1e-5, 1e-5, 'gradient field vs grid_sampler(sample_grad=True)')
    close(m.query_sdf_field(24, smooth=False), g['field_u_raw'], 1e-5, 2e-6)
    res = 37
    ref = R.sdf_field(om['sdf'], om['xyz_min'], om['xyz_max'], res, smooth=True, sigma=0.5)
    u = m.query_sdf_field(res, smooth=True, sigma=0.5)
    assert u.shape == (res, res, res)
    close(u, ref, 1e-5, 2e-6)
    sref, gref = R.sdf_gradient_field(om['sdf'], om['xyz_min'], om['xyz_max'], om['voxel_size'], res, smooth=True, sigma=0.5)
    u2, g2 = m.query_sdf_field(res, with_gradient=True)
    assert torch.equal(u2, u)
    close(g2, gref, 1e-5, 1e-5, 'gradient field')
    # identical to the chunked meshgrid -> grid_sampler path, bit for bit
    grid = m.mesh_query_grid(True, 0.5)
    ax = [torch.linspace(-1., 1., res) for _ in range(3)]
    pts = torch.stack(torch.meshgrid(*ax, indexing='ij'), -1).reshape(-1, 3).to(DEV)
    s_, gs_, _ = m.grid_sampler(pts, grid, sample_ret=True, sample_grad=True)
    assert torch.equal(-s_.reshape(res, res, res), u) and torch.equal(gs_.reshape(res, res, res, 3), g2)
    slabs = [m.query_sdf_field(res, x_range=(a, min(a + 10, res)), with_gradient=True, sdf_grid=grid) for a in range(0, res, 10)]
    assert torch.equal(torch.cat([s[0] for s in slabs], 0), u) and torch.equal(torch.cat([s[1] for s in slabs], 0), g2)
    close(m.query_sdf_field(res, smooth=False), R.sdf_field(om['sdf'], om['xyz_min'], om['xyz_max'], res, smooth=False), 1e-5, 2e-6)


@pytest.mark.parametrize('cl', [False, True])
def test_scale_volume_grid_matches_oracle(cl):
    """Progressive growing (lib/voxurf_fine.py:384-397, lib/grid.py:60-65): resampled grids, new voxel size, new
    non-empty mask and the sdf = 1 fill outside it."""
    sc = S.make_fine_scene(20, 6, 32, seed=3, mask_G=12)
    m = product_fine_model(sc, k0_channels_last=cl)
    om = oracle_fine_model(sc, requires_grad=False)
    new_voxels = 23000   # 28.4^3: clear of the float32 pow / floor edge at exact cubes (the golden vector uses the same)
    m.scale_volume_grid(new_voxels)
    ws = tuple(int(w) for w in m.world_size)
    assert ws == (28, 28, 28) and m.sdf.grid.shape == (1, 1) + ws and m.k0.grid.shape == (1, 6) + ws
    nonempty = R.nonempty_mask(om['mask_cache'], om['xyz_min'], om['xyz_max'], ws)
    assert torch.equal(m.nonempty_mask.cpu(), nonempty)
    sdf = R.scale_volume(om['sdf'], ws)
    sdf[~nonempty] = 1
    close(m.sdf.grid, sdf, 1e-5, 1e-6)
    close(m.k0.grid, R.scale_volume(om['k0'], ws), 1e-5, 1e-6)
    g = load_golden('r2_extras.npz')      # the reference's own scale_volume_grid (lib/voxurf_fine.py:384-397)
    assert tuple(int(w) for w in g['sv_world_size']) == ws and (m.nonempty_mask.cpu().numpy() == g['sv_nonempty']).all()
    close(m.sdf.grid, g['sv_sdf'], 1e-5, 1e-6, 'sdf vs reference'); close(m.k0.grid, g['sv_k0'], 1e-5, 1e-6, 'k0 vs reference')
    assert m.k0.grid.is_contiguous(memory_format=torch.channels_last_3d) == cl or not cl
    voxel_size = ((om['xyz_max'] - om['xyz_min']).prod() / new_voxels).pow(1 / 3)
    close(m.voxel_size, voxel_size, 1e-6, 0)
    # the resized model still renders
    ro, rd, vd = (T(x).to(DEV) for x in S.make_rays(64, seed=3))
    ret = m(ro, rd, vd, global_step=100, near=0.3, far=6.0, bg=0.0, stepsize=0.5)
    assert torch.isfinite(ret['rgb_marched']).all()


def test_coarse_regularisers_match_reference_formula():
    """lib/voxurf_coarse.py:300-320,702-715 (sum / 3 / mask.sum(), not the fine file's per-axis means): values and gradients of
    density_total_variation(sdf_tv, smooth_grad_tv) and k0_total_variation against the reference's own functions."""
    g = load_golden('r2_extras.npz')
    sc = S.make_coarse_scene(16, 12, 32, seed=4, mask_G=12)
    m = product_coarse_model(sc)
    m.gradient = m.neus_sdf_gradient(sdf=m.sdf.grid)
    tv = m.density_total_variation(sdf_tv=0.1, smooth_grad_tv=0.2)
    close(tv, g['ctv_density_value'], 1e-5, 1e-8)
    tv.backward()
    close(m.sdf.grid.grad, g['ctv_density_grad'], 1e-4, 1e-7)
    k0l = m.k0_total_variation()
    close(k0l, g['ctv_k0_value'], 1e-5, 1e-8)
    k0l.backward()
    close(m.k0.grid.grad, g['ctv_k0_grad'], 1e-4, 1e-9)
