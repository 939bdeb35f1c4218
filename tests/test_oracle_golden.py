"""CPU: the oracle (oracle/voxurf_ref.py + oracle/ref_kernels.c) against vectors produced by the
reference's own Python code (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import torch

from oracle import kernels as K
from oracle import voxurf_ref as R
from voxurf_b200 import synthetic as S
from tests.helpers import T, load_golden, oracle_coarse_model, oracle_fine_model

XYZ_MIN, XYZ_MAX = torch.tensor([-1., -1., -1.]), torch.tensor([1., 1., 1.])


def close(a, b, rtol=1e-5, atol=1e-6):
    a = a.detach().numpy() if torch.is_tensor(a) else np.asarray(a)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def _ops_inputs():
    rs = np.random.RandomState(11)
    pts = rs.uniform(-1.0, 1.0, (300, 3)).astype(np.float32)
    pts[:8] = np.array([[-1, -1, -1], [1, 1, 1], [1, -1, 0.3], [0.999, 0.2, -0.999], [0, 0, 0],
                        [-0.95, 0.95, 0.5], [0.5, 1.0, -1.0], [-1.0, 0.1, 0.2]], np.float32)
    return rs, T(pts)


def test_fine_ops_match_reference():
    g = load_golden('fine_ops.npz')
    sc = S.make_fine_scene(20, 6, 32, seed=3, mask_G=12)
    m = oracle_fine_model(sc, requires_grad=False, apply_nonempty=False)
    rs, pts = _ops_inputs()
    close(m['voxel_size'], g['voxel_size'], 0, 0)
    sdf, grad, feat = R.fine_grid_sampler(pts, m['sdf'], XYZ_MIN, XYZ_MAX, m['voxel_size'])
    close(sdf, g['gs_sdf']); close(grad, g['gs_grad']); close(feat, g['gs_feat'])
    f4, g4 = R.sample_sdfs(pts, m['sdf'], [0.5, 1.0, 1.5, 2.0], XYZ_MIN, XYZ_MAX, m['voxel_size'], use_grad_norm=True)
    close(f4, g['ss_feat']); close(g4, g['ss_grad'])
    _, g4r = R.sample_sdfs(pts, m['sdf'], [0.5, 1.0, 1.5, 2.0], XYZ_MIN, XYZ_MAX, m['voxel_size'], use_grad_norm=False)
    close(g4r, g['ss_grad_raw'])
    close(R.dense_grid_forward(m['k0'], pts, XYZ_MIN, XYZ_MAX), g['k0'])
    close(R.sdf_gradient_grid(m['sdf'], m['voxel_size']), g['fd_gradient'])
    close(R.gaussian_kernel3d(5, 0.8), g['smooth_k5_weight'][0, 0], 1e-6, 0)
    close(R.tv_smooth_kernel(), g['tv_smooth_weight'][0, 0], 1e-6, 0)
    close(R.conv3d_replicate(m['sdf'], R.gaussian_kernel3d(5, 0.8)), g['smooth_k5'])
    close(R.conv3d_replicate(m['sdf'], R.gaussian_kernel3d(3, 0.5)), g['smooth_k3'])
    mc = m['mask_cache']
    close(mc['density'], g['mask_cache_density'], 0, 0)
    out = R.mask_cache_forward(mc['density'], pts, mc['xyz_min'], mc['xyz_max'], mc['act_shift'], mc['voxel_size_ratio'], mc['thres'])
    assert (out.numpy() == g['mask_cache']).all()
    assert (m['nonempty_mask'].numpy() == g['nonempty_mask']).all()
    close(R.total_variation(m['sdf'], m['nonempty_mask']), g['tv_value'])
    # NeuS alpha
    n_rays = 16
    vd = T(S.make_rays(n_rays, seed=5)[2])
    rid = T(np.sort(rs.randint(0, n_rays, 300)).astype(np.int64))
    a_sdf = T((rs.standard_normal(300) * 0.05).astype(np.float32))
    a_grad = T(rs.standard_normal((300, 3)).astype(np.float32))
    s_val = R.s_val_schedule(1500, 50, 0.05)
    assert s_val == float(g['alpha_s_val'])
    held = float(torch.ones(1) * s_val)
    assert np.float32(held) == g['alpha_s_val_held'][0]
    close(R.neus_alpha_from_sdf_scatter(vd, rid, 0.5 * m['voxel_size'], a_sdf, a_grad, held), g['alpha'])
    # smooth-grad TV value and gradient
    sdf = m['sdf'].clone().requires_grad_(True)
    tv = R.smooth_grad_tv(R.sdf_gradient_grid(sdf, m['voxel_size']), m['nonempty_mask'], 0.05)
    tv.backward()
    close(tv, g['sgtv_value']); close(sdf.grad, g['sgtv_grad'], 1e-5, 1e-9)


def _check_ret(ret, g, keys):
    for k in keys:
        if k in g:
            a = ret[k]
            if a.dtype == torch.bool:
                assert (a.numpy() == g[k]).all(), k
            else:
                np.testing.assert_allclose(a.detach().numpy(), g[k], rtol=1e-5, atol=1e-6, err_msg=k)


def test_fine_forward_backward_match_reference():
    g = load_golden('fine_forward.npz')
    sc = S.make_fine_scene(20, 6, 32, seed=3, mask_G=12)
    m = oracle_fine_model(sc)
    close(m['sdf'], g['sdf_after_mask'], 0, 0)
    ro, rd, vd = (T(x) for x in S.make_rays(96, seed=777))
    target = T(S.make_target(vd.numpy()))
    ret = R.fine_forward(m, ro, rd, vd, 15001, near=0.3, stepsize=0.5, bg=0, render_grad=True, render_depth=True)
    assert ret['s_val'] == float(g['s_val'])
    _check_ret(ret, g, ['alphainv_cum', 'weights', 'rgb_marched', 'rgb_marched0', 'normal_marched', 'raw_alpha',
                        'raw_rgb', 'depth', 'disp', 'mask', 'mask_outbbox', 'gradient'])
    close(ret['_full_gradient'], g['full_gradient'])
    loss = R.fine_loss(ret, target)
    close(loss, g['loss'])
    loss.backward()
    close(m['sdf'].grad, g['grad_sdf'], 1e-4, 1e-8)
    close(m['k0'].grad, g['grad_k0'], 1e-4, 1e-8)
    for i, (W, b) in enumerate(m['rgbnet']):
        close(W.grad, g[f'grad_rgbnet_W{i}'], 1e-4, 1e-7); close(b.grad, g[f'grad_rgbnet_b{i}'], 1e-4, 1e-7)
    for i, (W, b) in enumerate(m['k_rgbnet']):
        close(W.grad, g[f'grad_k_rgbnet_W{i}'], 1e-4, 1e-7); close(b.grad, g[f'grad_k_rgbnet_b{i}'], 1e-4, 1e-7)
    # TV add-grad (total_variation_kernel.cu) then the trainer's Adam (utils.py:154-199)
    gs = m['sdf'].grad.clone().contiguous()
    w = 0.01 * 0.1 / 96 * 20 / 128
    K.total_variation_add_grad(m['sdf'].detach().contiguous(), gs, w, w, w, True)
    close(gs, g['grad_sdf_after_tv'], 1e-4, 1e-8)
    p = m['sdf'].detach().clone()
    R.python_adam_step(p, gs, torch.zeros_like(p), torch.zeros_like(p), 1, 5e-3)
    close(p, g['sdf_after_adam'], 1e-5, 1e-6)
    p2 = m['k0'].detach().clone()
    R.python_adam_step(p2, m['k0'].grad, torch.zeros_like(p2), torch.zeros_like(p2), 1, 1e-1)
    close(p2, g['k0_after_adam'], 1e-5, 1e-6)
    m['sdf'], m['k0'] = p, p2   # the eval-mode vectors were taken after the optimizer step
    with torch.no_grad():
        m['s_val'] = float(torch.ones(1) * R.s_val_schedule(15001, 50, 0.05))
        ret_e = R.fine_forward(m, ro, rd, vd, None, near=0.3, stepsize=0.5, bg=0, render_grad=True, render_depth=True)
    close(ret_e['rgb_marched'], g['eval_rgb_marched']); close(ret_e['normal_marched'], g['eval_normal_marched'])
    close(ret_e['depth'], g['eval_depth'])


def test_coarse_forward_backward_match_reference():
    g = load_golden('coarse_forward.npz')
    sc = S.make_coarse_scene(16, 12, 32, seed=4, mask_G=12)
    m = oracle_coarse_model(sc)
    close(m['sdf'], g['sdf_after_mask'], 0, 0)
    ro, rd, vd = (T(x) for x in S.make_rays(96, seed=777))
    target = T(S.make_target(vd.numpy()))
    ret = R.coarse_forward(m, ro, rd, vd, 2000, near=0.3, stepsize=0.5, bg=0, render_grad=True)
    _check_ret(ret, g, ['alphainv_cum', 'weights', 'rgb_marched', 'normal_marched', 'raw_alpha', 'raw_rgb', 'mask',
                        'mask_outbbox', 'gradient'])
    loss = torch.nn.functional.mse_loss(ret['rgb_marched'], target)
    close(loss, g['loss'])
    loss.backward()
    close(m['sdf'].grad, g['grad_sdf'], 1e-4, 1e-8)
    close(m['k0'].grad, g['grad_k0'], 1e-4, 1e-8)
    for i, (W, b) in enumerate(m['rgbnet']):
        close(W.grad, g[f'grad_rgbnet_W{i}'], 1e-4, 1e-7); close(b.grad, g[f'grad_rgbnet_b{i}'], 1e-4, 1e-7)
    # dense [N,S] formulation (BASELINE.json config 1)
    with torch.no_grad():
        pts, mask, step = R.sample_ray_ori(ro[:8], rd[:8], XYZ_MIN, XYZ_MAX, (16, 16, 16), 0.3, 1e9, 0.5, m['voxel_size'])
        close(pts, g['dense_pts']); assert (mask.numpy() == g['dense_mask']).all(); close(step, g['dense_step'])
        S_ = pts.shape[1]
        sdf = R.grid_trilinear(m['sdf'], pts.reshape(-1, 3), XYZ_MIN, XYZ_MAX).reshape(8, S_)
        grad = R.grid_trilinear(R.sdf_gradient_grid(m['sdf'], m['voxel_size']), pts.reshape(-1, 3), XYZ_MIN, XYZ_MAX).reshape(8, S_, 3)
        close(sdf, g['dense_sdf']); close(grad, g['dense_grad'])
        alpha = R.neus_alpha_from_sdf_dense(vd[:8], step, sdf, grad, float(torch.ones(1) * R.s_val_schedule(2000, 50, 0.2)))
        close(alpha, g['dense_alpha'])
        w, cum = R.get_ray_marching_ray(alpha)
        close(w, g['dense_weights']); close(cum, g['dense_alphainv_cum'])


def test_python_adam_matches_reference():
    g = load_golden('adam.npz')
    rs = np.random.RandomState(21)
    p = T(rs.standard_normal(257).astype(np.float32))
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for it in range(3):
        gr = rs.standard_normal(257).astype(np.float32)
        gr[::5] = 0
        R.python_adam_step(p, T(gr), m, v, it + 1, 5e-3)
        close(p, g['traj'][it], 1e-6, 1e-7)
    close(m, g['exp_avg'], 1e-6, 1e-8); close(v, g['exp_avg_sq'], 1e-6, 1e-10)


def test_maskgrid_matches_reference():
    g = load_golden('maskgrid.npz')
    rs = np.random.RandomState(21)
    rs.standard_normal(257)
    for _ in range(3):
        rs.standard_normal(257)
    mk = rs.uniform(0, 1, (9, 10, 11)) > 0.5
    scale, shift = R.mask_grid_params(mk.shape, XYZ_MIN, XYZ_MAX)
    close(scale, g['scale'], 0, 0); close(shift, g['shift'], 0, 0)
    q = T(rs.uniform(-1.2, 1.2, (400, 3)).astype(np.float32))
    assert (K.maskcache_lookup(T(mk), q, scale, shift).numpy() == g['out']).all()


def _train_views():
    views = [S.make_view(seed=40 + i, H=h, W=w, inverse_y=False) for i, (h, w) in enumerate(S.TRAIN_VIEW_SIZES)]
    imgs = [T(S.make_image(h, w, seed=i)) for i, (h, w, _, _) in enumerate(views)]
    return views, imgs


def test_view_rays_match_reference():
    """get_rays_of_a_view / ndc_rays (lib/voxurf_fine.py:1001-1070) as run by the reference's own code."""
    g = load_golden('rays.npz')
    for name, kw in S.RAY_CASES.items():
        H, W, K_, c2w = S.make_view(seed=kw['seed'], H=kw['H'], W=kw['W'], inverse_y=kw['inverse_y'])
        ro, rd, vd = R.view_rays(H, W, T(K_), T(c2w), ndc=kw['ndc'], inverse_y=kw['inverse_y'], flip_x=kw['flip_x'],
                                 flip_y=kw['flip_y'], mode=kw['mode'])
        close(ro, g[name + '_rays_o'], 2e-6, 1e-6); close(rd, g[name + '_rays_d'], 2e-6, 1e-6)
        close(vd, g[name + '_viewdirs'], 2e-6, 1e-7)


def test_training_rays_in_maskcache_match_reference():
    """get_training_rays_in_maskcache_sampling (lib/voxurf_fine.py:1127-1164): per-view counts and row order exact."""
    g = load_golden('rays.npz')
    sc = S.make_fine_scene(20, 6, 32, seed=3, mask_G=12)
    m = oracle_fine_model(sc, requires_grad=False)
    views, imgs = _train_views()
    rgb, ro, rd, vd, counts = R.training_rays_in_maskcache(
        imgs, [T(v[3]) for v in views], [(v[0], v[1]) for v in views], [T(v[2]) for v in views], m['mask_cache'], XYZ_MIN,
        XYZ_MAX, m['voxel_size'], near=0.3, stepsize=0.5)
    assert counts == g['tr_imsz'].tolist()
    assert np.array_equal(rgb.numpy(), g['tr_rgb'])          # the kept pixels, in order: identifies the kept rays exactly
    close(ro, g['tr_rays_o'], 2e-6, 1e-6); close(rd, g['tr_rays_d'], 2e-6, 1e-6); close(vd, g['tr_viewdirs'], 2e-6, 1e-7)


# ------------------------------------------------------------------------------------------------
# round 2: tests/golden/r2_extras.npz (make_golden_r2.py) -- coarse regularisers, progressive growing, mesh fields,
# vertex colours, all from the reference's own functions
# ------------------------------------------------------------------------------------------------
def test_coarse_total_variation_is_not_the_fine_formula():
    """lib/voxurf_coarse.py:702-715 divides the three masked sums by 3 * mask.sum(); lib/voxurf_fine.py:956-969 takes
    per-axis means.  The two differ whenever the mask has a boundary."""
    g = load_golden('r2_extras.npz')
    sc = S.make_coarse_scene(16, 12, 32, seed=4, mask_G=12)
    m = oracle_coarse_model(sc)
    close(R.total_variation_coarse(m['sdf'], m['nonempty_mask']), g['ctv_value'])
    assert abs(float(R.total_variation(m['sdf'], m['nonempty_mask'])) - float(g['ctv_value'])) > 1e-3 * float(g['ctv_value'])
    G = R.sdf_gradient_grid(m['sdf'], m['voxel_size'])
    tv = R.total_variation_coarse(m['sdf'], m['nonempty_mask']) / 2 / m['voxel_size'] * 0.1 \
        + R.smooth_grad_tv(G, m['nonempty_mask'], 0.2)
    close(tv, g['ctv_density_value'])
    tv.backward()
    close(m['sdf'].grad, g['ctv_density_grad'], 1e-5, 1e-7)
    k0l = R.total_variation_coarse(m['k0'], m['nonempty_mask'].repeat(1, 12, 1, 1, 1))
    close(k0l, g['ctv_k0_value'])
    k0l.backward()
    close(m['k0'].grad, g['ctv_k0_grad'], 1e-5, 1e-9)


def test_scale_volume_matches_reference():
    g = load_golden('r2_extras.npz')
    sc = S.make_fine_scene(20, 6, 32, seed=3, mask_G=12)
    m = oracle_fine_model(sc, requires_grad=False)
    ws = tuple(int(w) for w in g['sv_world_size'])
    assert ws == (28, 28, 28)
    nonempty = R.nonempty_mask(m['mask_cache'], XYZ_MIN, XYZ_MAX, ws)
    assert (nonempty.numpy() == g['sv_nonempty']).all()
    sdf = R.scale_volume(m['sdf'], ws)
    sdf[~nonempty] = 1
    close(sdf, g['sv_sdf'])
    close(R.scale_volume(m['k0'], ws), g['sv_k0'])


def test_mesh_fields_and_vertex_colours_match_reference():
    g = load_golden('r2_extras.npz')
    sc = S.make_fine_scene(24, 6, 32, seed=5, mask_G=12)
    m = oracle_fine_model(sc, requires_grad=False)
    close(R.sdf_field(m['sdf'], XYZ_MIN, XYZ_MAX, 24, smooth=True, sigma=0.5), g['field_u'])
    close(R.sdf_field(m['sdf'], XYZ_MIN, XYZ_MAX, 24, smooth=False), g['field_u_raw'])
    s, gr = R.sdf_gradient_field(m['sdf'], XYZ_MIN, XYZ_MAX, m['voxel_size'], 24, smooth=True, sigma=0.5)
    close(s, g['field_sdf']); close(gr, g['field_grad'], 1e-5, 1e-5)
    close(R.mesh_color_forward(m, T(g['mc_pts'])), g['mc_rgb'], 1e-5, 2e-6)
