"""GPU parity, operator by operator: the CUDA path through the C ABI vs the CPU oracle on the same
seeded inputs.  Bars: integer / index / bool outputs bit-exact; fp32 forward rtol 1e-5, gradients rtol 1e-4
(north_star), with small absolute floors written next to each check."""
import numpy as np
import pytest
import torch

from oracle import kernels as K
from oracle import voxurf_ref as R
from voxurf_b200 import synthetic as S
from tests.helpers import T

pytestmark = pytest.mark.gpu

XYZ_MIN, XYZ_MAX = torch.tensor([-1., -1., -1.]), torch.tensor([1., 1., 1.])
MN, MX = [-1., -1., -1.], [1., 1., 1.]
DEV = 'cuda'


def cu(x):
    return x.to(DEV)


def close(a, b, rtol=1e-5, atol=1e-6, msg=''):
    np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy(), rtol=rtol, atol=atol, err_msg=msg)


def exact(a, b, msg=''):
    a, b = a.detach().cpu(), b.detach().cpu()
    assert a.shape == b.shape and a.dtype == b.dtype, (msg, a.shape, b.shape, a.dtype, b.dtype)
    assert torch.equal(a, b), (msg, (a != b).sum().item())


def rays(n, seed=777, special=True):
    o, d, v = (T(x) for x in S.make_rays(n, seed=seed))
    if special and n >= 8:
        d[0, 1] = 0.0                                   # zero direction component -> 1e-6 (hazard 1)
        d[1] = torch.tensor([0.0, 0.0, 1.0]); o[1] = torch.tensor([0.2, -0.3, -3.0])   # axis aligned
        o[2] = torch.tensor([5.0, 5.0, 5.0]); d[2] = torch.tensor([1.0, 0.2, 0.1])     # misses the box (hazard 2)
        o[3] = torch.tensor([0.1, 0.1, 0.1])                                           # origin inside the box
        o[4] = torch.tensor([-3.0, 1.0, 0.0]); d[4] = torch.tensor([1.0, 0.0, 0.0])    # grazes a face (hazard 3)
    return o, d, v


# ------------------------------------------------------------------------------------------------ sampling
@pytest.mark.parametrize('n,G,near', [(257, 32, 0.3), (8192, 256, 0.3), (1, 16, 0.05), (64, 96, 2.0)])
def test_sampling_bit_exact(n, G, near):
    from voxurf_b200 import render_utils_cuda as ru
    o, d, _ = rays(n, special=n >= 8)
    stepdist = float(np.float32(0.5 * 2.0 / G))
    ref = K.sample_pts_on_rays(o, d, XYZ_MIN, XYZ_MAX, near, 1e9, stepdist)
    t_min, t_max = ru.infer_t_minmax(cu(o), cu(d), cu(XYZ_MIN), cu(XYZ_MAX), near, 1e9)
    exact(t_min, ref[5], 't_min'); exact(t_max, ref[6], 't_max')
    exact(ru.infer_n_samples(cu(d), t_min, t_max, stepdist), ref[4], 'n_samples')
    s, dr = ru.infer_ray_start_dir(cu(o), cu(d), t_min)
    rs, rd = K.infer_ray_start_dir(o, d, ref[5])
    exact(s, rs, 'rays_start'); exact(dr, rd, 'rays_dir')
    out = ru.sample_pts_on_rays(cu(o), cu(d), cu(XYZ_MIN), cu(XYZ_MAX), near, 1e9, stepdist)
    for a, b, name in zip(out, ref, ['pts', 'mask_outbbox', 'ray_id', 'step_id', 'N_steps', 't_min', 't_max']):
        exact(a, b, name)


def test_sampling_empty_and_misc():
    from voxurf_b200 import render_utils_cuda as ru
    e = torch.zeros(0, 3, device=DEV)
    out = ru.sample_pts_on_rays(e, e, cu(XYZ_MIN), cu(XYZ_MAX), 0.1, 1e9, 0.01)
    assert out[0].shape == (0, 3) and out[2].shape == (0,) and out[4].shape == (0,)
    o, d, _ = rays(33)
    p, m = ru.sample_ndc_pts_on_rays(cu(o), cu(d), cu(XYZ_MIN), cu(XYZ_MAX), 17)
    rp, rm = K.sample_ndc_pts_on_rays(o, d, XYZ_MIN, XYZ_MAX, 17)
    exact(p, rp); exact(m, rm)
    tmax = torch.rand(33) + 3
    close(ru.sample_bg_pts_on_rays(cu(o), cu(d), cu(tmax), 0.3, 19), K.sample_bg_pts_on_rays(o, d, tmax, 0.3, 19), 1e-6, 1e-7)


def test_maskcache_lookup_bit_exact():
    from voxurf_b200 import render_utils_cuda as ru
    rs = np.random.RandomState(3)
    world = T(rs.uniform(0, 1, (9, 10, 11)) > 0.5)
    q = T(rs.uniform(-1.2, 1.2, (5000, 3)).astype(np.float32))
    q[:4] = torch.tensor([[-1., -1., -1.], [1., 1., 1.], [0.0, 0.1, -0.1], [-1.1, 0., 0.]])   # .5 rounding, OOB
    scale, shift = R.mask_grid_params(world.shape, XYZ_MIN, XYZ_MAX)
    exact(ru.maskcache_lookup(cu(world), cu(q), cu(scale), cu(shift)), K.maskcache_lookup(world, q, scale, shift))
    assert ru.maskcache_lookup(cu(world), torch.zeros(0, 3, device=DEV), cu(scale), cu(shift)).shape == (0,)


def test_fused_march_matches_staged_reference_path():
    """vx_march_* == sample_pts_on_rays -> in-bbox filter -> MaskCache -> filter (voxurf_fine.py:593-636)."""
    from tests.helpers import product_fine_model, oracle_fine_model
    sc = S.make_fine_scene(48, 6, 16, seed=5, mask_G=20)
    m = product_fine_model(sc, apply_nonempty=False)
    om = oracle_fine_model(sc, requires_grad=False, apply_nonempty=False)
    o, d, _ = rays(700)
    out = m._march(cu(o), cu(d), 0.3, 0.5)
    pts, rid, sid, mob, _ = R.sample_ray(o, d, XYZ_MIN, XYZ_MAX, 0.3, 0.5, om['voxel_size'])
    mc = om['mask_cache']
    keep = R.mask_cache_forward(mc['density'], pts, mc['xyz_min'], mc['xyz_max'], mc['act_shift'], mc['voxel_size_ratio'], mc['thres'])
    mob[~mob.clone()] |= ~keep
    exact(out['ray_id'].long(), rid[keep], 'ray_id'); exact(out['step_id'].long(), sid[keep], 'step_id')
    exact(out['ray_pts'], pts[keep], 'ray_pts'); exact(out['mask_outbbox'], mob, 'mask_outbbox')
    # standalone MaskCache.forward and hit_coarse_geo
    exact(m.mask_cache(cu(pts)), keep, 'mask_cache')
    hit = torch.zeros(700, dtype=torch.bool); hit[rid[keep]] = True
    exact(m.hit_coarse_geo(cu(o), cu(d), 0.3, 6.0, 0.5), hit, 'hit')


@pytest.mark.parametrize('kind', ['smooth', 'noise', 'on_threshold'])
def test_march_cell_verdicts_are_exact(kind):
    """vx_march_flags_cells (per-cell pass / fail / evaluate verdicts, vx_mask_cache_cells) sets exactly the keep bits of
    vx_march_flags (every sample through the 8-corner interpolation + softplus + exp of lib/voxurf_fine.py:930-942)."""
    from voxurf_b200._lib import call
    rs = np.random.RandomState(17)
    G, N = 40, 3000
    act_shift, ratio, thres = -4.0, 0.5, 1e-3
    ax = np.linspace(-1, 1, G, dtype=np.float32)
    r = np.sqrt(ax[:, None, None] ** 2 + ax[None, :, None] ** 2 + ax[None, None, :] ** 2)
    if kind == 'smooth':
        dens = (0.6 - r) * 30                                   # a ball of occupied space, free outside
    elif kind == 'noise':
        dens = rs.standard_normal((G, G, G)) * 6                # verdicts change from cell to cell
    else:
        # constant density sitting (in fp32) on the threshold itself, + noise of a few ulp: no cell may get a verdict
        d0 = float(np.log(np.expm1(-np.log1p(-thres) / ratio))) - act_shift
        dens = np.full((G, G, G), d0) * (1 + rs.randint(-3, 4, (G, G, G)) * 6e-8)
    dens = cu(T(dens.astype(np.float32))).contiguous()
    cells = torch.empty(G, G, G, dtype=torch.uint8, device=DEV)
    call('vx_mask_cache_cells', dens, G, G, G, act_shift, ratio, thres, cells)
    hist = torch.bincount(cells.view(-1).long(), minlength=3).tolist()
    if kind == 'smooth':
        assert hist[0] > 0.3 * G ** 3 and hist[1] > 0.05 * G ** 3 and 0 < hist[2] < 0.25 * G ** 3, hist
    if kind == 'on_threshold':
        assert hist[0] == 0 and hist[1] == 0, hist
    o, d, _ = rays(N)
    o, d = cu(o), cu(d)
    mn, mx = cu(XYZ_MIN), cu(XYZ_MAX)
    stepdist = 0.5 * 2.0 / 96
    t_min, t_max = torch.empty(N, device=DEV), torch.empty(N, device=DEV)
    n_steps = torch.empty(N, dtype=torch.int64, device=DEV)
    start, dirs = torch.empty(N, 3, device=DEV), torch.empty(N, 3, device=DEV)
    offsets = torch.empty(N + 1, dtype=torch.int64, device=DEV)
    call('vx_ray_setup', o, d, mn, mx, 0.3, 1e9, stepdist, N, t_min, t_max, n_steps, start, dirs, offsets)
    words = int(offsets[-1]) // 32 + N + 1
    out = []
    for tab in (None, cells):
        bi, bk = (torch.zeros(words, dtype=torch.int32, device=DEV) for _ in range(2))
        kc, ko = torch.zeros(N, dtype=torch.int32, device=DEV), torch.zeros(N + 1, dtype=torch.int32, device=DEV)
        # the mask grid spans a smaller box than the rays' box: samples outside it take the exact (zero-padded) path
        call('vx_march_flags_cells', start, dirs, mn, mx, offsets, N, stepdist, dens, G, G, G, [-0.9, -0.9, -0.9], [0.9, 0.9, 0.9],
             act_shift, ratio, thres, tab, bi, bk, kc, ko)
        out.append((bi, bk, kc, ko))
    for a, b in zip(*out):
        assert torch.equal(a, b)
    assert int(out[0][3][-1]) > 0 or kind == 'on_threshold'


# ------------------------------------------------------------------------------------------------ raw2alpha / alpha2weight
def test_raw2alpha():
    from voxurf_b200 import render_utils_cuda as ru
    torch.manual_seed(5)
    d = torch.randn(10000) * 8
    d[:3] = torch.tensor([100., -100., 0.])
    e, a = ru.raw2alpha(cu(d), -4.0, 0.5)
    re, ra = K.raw2alpha(d, -4.0, 0.5)
    # alpha = 1 - pow(1 + e, -interval): near alpha = 0 the result is quantised in steps of 2^-24 ~ 6e-8 and CUDA's powf
    # (<= 2 ulp) and glibc's (< 1 ulp) may land on neighbouring steps -> absolute floor of 4 steps
    close(e, re, 2e-6, 0); close(a, ra, 1e-5, 2.5e-7)
    gb = torch.randn(10000)
    close(ru.raw2alpha_backward(e, cu(gb), 0.5), K.raw2alpha_backward(e.cpu(), gb, 0.5), 1e-5, 1e-9)
    iv = torch.rand(10000) + 0.1
    e2, a2 = ru.raw2alpha_nonuni(cu(d), -4.0, cu(iv))
    close(a2, K.raw2alpha(d, -4.0, iv)[1], 1e-5, 2.5e-7)
    close(ru.raw2alpha_nonuni_backward(e2, cu(gb), cu(iv)), K.raw2alpha_backward(e2.cpu(), gb, iv), 1e-5, 1e-9)
    assert ru.raw2alpha(torch.zeros(0, device=DEV), 0.0, 0.5)[1].shape == (0,)


def _ragged_ray_ids(n_rays, rs, max_len=300, empty_frac=0.2):
    lens = rs.randint(0, max_len, n_rays)
    lens[rs.uniform(0, 1, n_rays) < empty_frac] = 0
    return torch.from_numpy(np.repeat(np.arange(n_rays), lens).astype(np.int64))


@pytest.mark.parametrize('n_rays,scale', [(64, 0.05), (2048, 0.3), (8192, 0.02)])
def test_alpha2weight_bit_exact(n_rays, scale):
    from voxurf_b200 import render_utils_cuda as ru
    rs = np.random.RandomState(n_rays)
    rid = _ragged_ray_ids(n_rays, rs)
    alpha = T((rs.uniform(0, 1, rid.shape[0]) ** 3 * scale * 3).clip(0, 1).astype(np.float32))
    alpha[::97] = 1.0; alpha[::89] = 0.0
    out = ru.alpha2weight(cu(alpha), cu(rid), n_rays)
    ref = K.alpha2weight(alpha, rid, n_rays)
    for a, b, name in zip(out, ref, ['weight', 'T', 'alphainv_last', 'i_start', 'i_end']):
        exact(a, b, name)
    gw = T(rs.standard_normal(rid.shape[0]).astype(np.float32)); gl = T(rs.standard_normal(n_rays).astype(np.float32))
    g = ru.alpha2weight_backward(cu(alpha), out[0], out[1], out[2], out[3], out[4], n_rays, cu(gw), cu(gl))
    exact(g, K.alpha2weight_backward(alpha, *ref, n_rays, gw, gl), 'grad_alpha')


def test_alpha2weight_empty():
    from voxurf_b200 import render_utils_cuda as ru
    out = ru.alpha2weight(torch.zeros(0, device=DEV), torch.zeros(0, dtype=torch.int64, device=DEV), 5)
    assert out[0].shape == (0,) and (out[2] == 1).all() and (out[3] == 0).all() and (out[4] == 0).all()


# ------------------------------------------------------------------------------------------------ TV / Adam
@pytest.mark.parametrize('dense', [True, False])
@pytest.mark.parametrize('masked', [False, True])
def test_total_variation_add_grad(dense, masked):
    from voxurf_b200 import total_variation_cuda as tv
    rs = np.random.RandomState(5)
    p = T(rs.standard_normal((1, 3, 13, 11, 17)).astype(np.float32) * 2)
    g = T(rs.standard_normal((1, 3, 13, 11, 17)).astype(np.float32)); g[rs.uniform(0, 1, g.shape) < 0.5] = 0
    mk = T((rs.uniform(0, 1, p.shape) > 0.3).astype(np.float32))
    gg, gr = cu(g.clone()), g.clone()
    if masked:
        tv.total_variation_add_grad_new(cu(p), gg, cu(mk), 0.3, 0.5, 0.7, dense)
        K.total_variation_add_grad(p, gr, 0.3, 0.5, 0.7, dense, mask=mk)
    else:
        tv.total_variation_add_grad(cu(p), gg, 0.3, 0.5, 0.7, dense)
        K.total_variation_add_grad(p, gr, 0.3, 0.5, 0.7, dense)
    close(gg, gr, 1e-6, 1e-7)


@pytest.mark.parametrize('mode', [0, 1, 2])
def test_adam_upd_reference_cuda_semantics(mode):
    from voxurf_b200 import adam_upd_cuda as ad
    rs = np.random.RandomState(7)
    n = 100003
    p, m, v = (T(rs.standard_normal(n).astype(np.float32)) for _ in range(3))
    v = v.abs()
    perlr = T(rs.uniform(0, 1, n).astype(np.float32))
    pc, mc, vc = cu(p.clone()), cu(m.clone()), cu(v.clone())
    for step in (1, 2, 30):
        g = T(rs.standard_normal(n).astype(np.float32)); g[::3] = 0
        if mode == 0:
            ad.adam_upd(pc, cu(g), mc, vc, step, 0.9, 0.99, 0.1, 1e-8)
        elif mode == 1:
            ad.masked_adam_upd(pc, cu(g), mc, vc, step, 0.9, 0.99, 0.1, 1e-8)
        else:
            ad.adam_upd_with_perlr(pc, cu(g), mc, vc, cu(perlr), step, 0.9, 0.99, 0.1, 1e-8)
        K.adam_upd(p, g, m, v, step, 0.9, 0.99, 0.1, 1e-8, mode=mode, perlr=perlr if mode == 2 else None)
        close(pc, p, 1e-6, 1e-6); close(mc, m, 1e-6, 1e-7); close(vc, v, 1e-6, 1e-8)


def test_trainer_adam_semantics_and_fused_zero_grad():
    from voxurf_b200.optim import Adam
    rs = np.random.RandomState(9)
    p0 = T(rs.standard_normal((1, 6, 9, 10, 11)).astype(np.float32))
    for cl in (False, True):
        p = torch.nn.Parameter(cu(p0.clone()).contiguous(memory_format=torch.channels_last_3d) if cl else cu(p0.clone()))
        opt = Adam([{'params': [p], 'lr': 0.1}], betas=(0.9, 0.99), zero_grad_in_step=True)
        rp, rm, rv = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
        rs2 = np.random.RandomState(10)
        for it in range(3):
            g = T(rs2.standard_normal(p0.shape).astype(np.float32)); g[..., ::2] = 0
            p.grad = cu(g.clone()).contiguous(memory_format=torch.channels_last_3d) if cl else cu(g.clone())
            opt.step()
            R.python_adam_step(rp, g, rm, rv, it + 1, 0.1)
            close(p.data, rp, 1e-6, 1e-7)
            assert (p.grad == 0).all()


@pytest.mark.parametrize('C', [12, 6, 8, 3])
def test_adam_touched_live_bitmaps_are_bit_identical_to_dense(C):
    """vx_adam_step with the touched / live voxel bitmaps (the fused step's k0 update) == the dense pass, bit for bit,
    over several steps in which the touched set moves; vx_bitmap_merge maintains `live`."""
    from voxurf_b200._lib import call
    rs = np.random.RandomState(31 + C)
    V = 4099 if C % 4 == 0 else (4098 if C % 2 == 0 else 4100)    # numel % 4 == 0; C = 6 makes float4 groups straddle two voxels
    N = V * C
    p0 = cu(T(rs.standard_normal(N).astype(np.float32)))
    state = {k: [p0.clone(), torch.zeros(N, device='cuda'), torch.zeros(N, device='cuda'), torch.zeros(N, device='cuda')]
             for k in ('dense', 'sparse', 'worklist')}
    n_words = (V + 31) // 32
    touched = torch.zeros(n_words, dtype=torch.int32, device='cuda')
    live = torch.zeros(n_words, dtype=torch.int32, device='cuda')
    touched_w, live_w = torch.zeros_like(touched), torch.zeros_like(live)     # the work-list form keeps its own pair
    work = torch.zeros(V + 1, dtype=torch.int32, device='cuda')
    ever = np.zeros(V, bool)
    for it in range(5):
        vox = rs.choice(V, size=V // 7, replace=False)
        ever[vox] = True
        g = np.zeros((V, C), np.float32)
        g[vox] = rs.standard_normal((vox.size, C)).astype(np.float32)
        g[vox[:5], 0] = 0      # zeros inside a touched voxel are still dense-updated
        bits = np.zeros(n_words * 32, bool); bits[vox] = True
        words = np.packbits(bits.reshape(-1, 32), axis=1, bitorder='little').view(np.uint32).reshape(-1).astype(np.int64)
        touched.copy_(torch.from_numpy(np.where(words >= 2 ** 31, words - 2 ** 32, words)).to(torch.int32))
        bc1, bc2 = 1 - 0.9 ** (it + 1), 1 - 0.99 ** (it + 1)
        touched_w.copy_(touched)
        for k, st in state.items():
            st[1].copy_(cu(T(g.reshape(-1))))
            if k == 'worklist':     # vx_adam_step_worklist: compacted word list, vx_bitmap_merge folded in
                call('vx_adam_step_worklist', st[0], st[1], st[2], st[3], N, 0.9, 0.99, 1 - 0.9, 1 - 0.99, 0.1 / bc1,
                     float(np.sqrt(bc2)), 1e-8, 1, touched_w, live_w, C, 1, work, None)
                continue
            bm = (touched, live) if k == 'sparse' else (None, None)
            call('vx_adam_step', st[0], st[1], st[2], st[3], None, N, 0.9, 0.99, 1 - 0.9, 1 - 0.99, 0.1 / bc1,
                 float(np.sqrt(bc2)), 1e-8, 0, 1, bm[0], bm[1], C, None)
        call('vx_bitmap_merge', live, touched, n_words)
        assert (touched == 0).all()
        got = np.unpackbits(live.cpu().numpy().view(np.uint8), bitorder='little')[:V].astype(bool)
        assert (got == ever).all()
        assert (touched_w == 0).all() and torch.equal(live_w, live)
        for a, b, c in zip(state['dense'], state['sparse'], state['worklist']):
            assert torch.equal(a, b) and torch.equal(a, c)
        assert (state['sparse'][1] == 0).all() and (state['worklist'][1] == 0).all()
    assert not ever.all()       # the skip path was exercised


# ------------------------------------------------------------------------------------------------ grid gathers
def _pts(n, rs, lo=-1.05, hi=1.05):
    p = T(rs.uniform(lo, hi, (n, 3)).astype(np.float32))
    p[:6] = torch.tensor([[-1., -1., -1.], [1., 1., 1.], [1., -1., 0.3], [0.999, 0.2, -0.999], [0., 0., 0.], [-1.0, 0.1, 0.2]])
    return p


@pytest.mark.parametrize('C,cl', [(1, False), (3, False), (6, False), (6, True), (12, True), (5, True), (4, True)])
def test_grid_gather_forward_backward(C, cl):
    from voxurf_b200 import ops
    rs = np.random.RandomState(C)
    grid = T(rs.standard_normal((1, C, 9, 12, 7)).astype(np.float32))
    pts = _pts(3000, rs)
    go = T(rs.standard_normal((3000, C)).astype(np.float32)); go[::4] = 0
    gr = grid.clone().requires_grad_(True)
    ref = R.grid_trilinear(gr, pts, XYZ_MIN, XYZ_MAX)
    ref.backward(go)
    gg = cu(grid.clone())
    if cl:
        gg = gg.contiguous(memory_format=torch.channels_last_3d)
    gg.requires_grad_(True)
    out = ops.grid_gather(gg, cu(pts), MN, MX)
    close(out, ref, 1e-5, 1e-6)
    out.backward(cu(go))
    close(gg.grad, gr.grad, 1e-4, 1e-5)


@pytest.mark.parametrize('disp,norm,xyz_order', [([1.0], False, True), ([0.5, 1.0, 1.5, 2.0], True, False),
                                                 ([0.5, 1.0, 1.5, 2.0], False, False), ([2.0], True, False)])
def test_sdf_taps_forward_backward(disp, norm, xyz_order):
    from voxurf_b200 import ops
    rs = np.random.RandomState(len(disp))
    G = 24
    vs = float(2.0 / G)
    grid = T(S.sphere_sdf(G, noise=0.05, seed=2))
    pts = _pts(2500, rs, -1.0, 1.0)
    gr = grid.clone().requires_grad_(True)
    L = len(disp)
    if xyz_order:
        sdf_r, grad_r, feat_r = R.fine_grid_sampler(pts, gr, XYZ_MIN, XYZ_MAX, torch.tensor(vs))
    else:
        feat_r, grad_r = R.sample_sdfs(pts, gr, disp, XYZ_MIN, XYZ_MAX, torch.tensor(vs), use_grad_norm=norm)
        sdf_r = R.grid_trilinear(gr, pts, XYZ_MIN, XYZ_MAX).squeeze(-1)
    g1, g2, g3 = (T(rs.standard_normal(t.shape).astype(np.float32)) for t in (sdf_r, feat_r, grad_r))
    g3[::5] = 0
    (sdf_r * g1).sum().backward(retain_graph=True); (feat_r * g2).sum().backward(retain_graph=True); (grad_r * g3).sum().backward()
    gg = cu(grid.clone()).requires_grad_(True)
    sdf, feat, grad = ops.sdf_taps(gg, cu(pts), MN, MX, disp, vs, use_grad_norm=norm, xyz_order=xyz_order, want_sdf=True)
    finite = torch.isfinite(grad_r).all(-1)
    close(sdf, sdf_r, 1e-5, 1e-6); close(feat, feat_r, 1e-5, 1e-6)
    close(grad[cu(finite)], grad_r[finite], 1e-4 if norm else 2e-5, 1e-5 if norm else 2e-5)
    ((sdf * cu(g1)).sum() + (feat * cu(g2)).sum() + (grad * cu(g3)).sum()).backward()
    close(gg.grad, gr.grad, 1e-4, 2e-4 * float(gr.grad.abs().max()))


def test_neus_alpha_forward_backward():
    from voxurf_b200 import ops
    rs = np.random.RandomState(4)
    n, n_rays = 20000, 64
    vd = T(S.make_rays(n_rays, seed=5)[2])
    rid = T(np.sort(rs.randint(0, n_rays, n)).astype(np.int64))
    sdf = T((rs.standard_normal(n) * 0.02).astype(np.float32)).requires_grad_(True)
    grad = T(rs.standard_normal((n, 3)).astype(np.float32)).requires_grad_(True)
    s_val = float(torch.ones(1) * R.s_val_schedule(15001, 50, 0.05))
    dist = float(np.float32(0.5 * 2 / 256))
    ref = R.neus_alpha_from_sdf_scatter(vd, rid, dist, sdf, grad, s_val)
    ga = T(rs.standard_normal(n).astype(np.float32))
    ref.backward(ga)
    s2, g2 = cu(sdf.detach()).requires_grad_(True), cu(grad.detach()).requires_grad_(True)
    for ids in (cu(rid), cu(rid).int()):
        s2.grad = g2.grad = None
        out = ops.neus_alpha(cu(vd), ids, s2, g2, dist, float(np.float32(1.0) / np.float32(s_val)))
        close(out, ref, 1e-5, 2e-6)
        out.backward(cu(ga))
        scale = float(sdf.grad.abs().max())
        close(s2.grad, sdf.grad, 1e-4, 1e-4 * scale); close(g2.grad, grad.grad, 1e-4, 1e-4 * float(grad.grad.abs().max()))


def test_segment_coo():
    from voxurf_b200.torch_scatter import segment_coo
    rs = np.random.RandomState(6)
    rid = _ragged_ray_ids(500, rs, max_len=40)
    for K_ in (1, 3, 7):
        src = T(rs.standard_normal((rid.shape[0], K_)).astype(np.float32)).requires_grad_(True)
        ref = R.segment_coo(src, rid, torch.zeros(500, K_))
        go = T(rs.standard_normal((500, K_)).astype(np.float32))
        ref.backward(go)
        s2 = cu(src.detach()).requires_grad_(True)
        out = segment_coo(s2, cu(rid), out=torch.zeros(500, K_, device=DEV), reduce='sum')
        close(out, ref, 1e-6, 1e-6)
        out.backward(cu(go))
        exact(s2.grad, src.grad)


# ------------------------------------------------------------------------------------------------ stencils
def test_fd_gradient_and_backward():
    from voxurf_b200 import ops
    for shape in [(11, 9, 13), (3, 3, 3), (2, 5, 4), (6, 5, 8), (5, 7, 12), (1, 1, 4)]:   # Z % 4 == 0: vectorised kernels
        rs = np.random.RandomState(sum(shape))
        sdf = T(rs.standard_normal((1, 1) + shape).astype(np.float32)).requires_grad_(True)
        vs = torch.tensor(0.0625)
        ref = R.sdf_gradient_grid(sdf, vs)
        go = T(rs.standard_normal(ref.shape).astype(np.float32))
        ref.backward(go)
        s2 = cu(sdf.detach()).requires_grad_(True)
        out = ops.fd_gradient(s2, float(vs))
        close(out, ref, 1e-6, 1e-6)
        out.backward(cu(go))
        close(s2.grad, sdf.grad, 1e-5, 1e-5)


@pytest.mark.parametrize('k,sigma', [(5, 0.8), (3, 0.5), (5, 1.0)])
def test_gaussian_conv_and_backward(k, sigma):
    from voxurf_b200 import ops
    from voxurf_b200.voxurf_fine import SmoothConv
    for shape in [(12, 10, 14), (5, 4, 6), (2, 3, 2)]:
        rs = np.random.RandomState(k)
        x = T(rs.standard_normal((1, 1) + shape).astype(np.float32)).requires_grad_(True)
        ref = R.conv3d_replicate(x, R.gaussian_kernel3d(k, sigma))
        go = T(rs.standard_normal(ref.shape).astype(np.float32))
        ref.backward(go)
        x2 = cu(x.detach()).requires_grad_(True)
        out = SmoothConv(k, sigma)(x2)
        close(out, ref, 1e-5, 1e-6)
        out.backward(cu(go))
        close(x2.grad, x.grad, 1e-4, 1e-5)


@pytest.mark.parametrize('shape', [(14, 12, 10), (9, 7, 12), (3, 2, 4)])
def test_smooth_grad_tv_value_and_gradient(shape):
    from voxurf_b200 import ops
    from voxurf_b200.voxurf_fine import _binomial_weights
    rs = np.random.RandomState(8)
    sdf = T(rs.standard_normal((1, 1) + shape).astype(np.float32)).requires_grad_(True)
    mask = T(rs.uniform(0, 1, (1, 1) + shape) > 0.4)
    vs = torch.tensor(0.1)
    ref = R.smooth_grad_tv(R.sdf_gradient_grid(sdf, vs), mask, 0.05)
    ref.backward()
    s2 = cu(sdf.detach()).requires_grad_(True)
    out = ops.smooth_grad_tv(ops.fd_gradient(s2, float(vs)), cu(mask[0, 0]), _binomial_weights(), 0.05, int(mask.sum()))
    close(out, ref, 1e-5, 1e-8)
    (out * 1.0).backward()
    close(s2.grad, sdf.grad, 1e-4, 1e-5 * float(sdf.grad.abs().max()))


@pytest.mark.parametrize('shape', [(9, 7, 12), (13, 11, 17), (2, 3, 4), (40, 33, 48)])
def test_sdf_regularisers_backward_and_vector_paths_are_bit_identical(shape):
    """The z-vectorised stencil kernels (Z % 4 == 0) and the fused vx_sdf_regularisers_backward produce exactly what
    the scalar kernels do (which the tests above / test_gpu_vs_reference.py pin to the oracle and the reference):
    compare a Z % 4 == 0 grid against the same data embedded in a grid that forces the scalar path (misaligned view)."""
    from voxurf_b200._lib import call
    from voxurf_b200 import total_variation_cuda as tv
    X, Y, Z = shape
    V = X * Y * Z
    rs = np.random.RandomState(V)
    sdf = cu(T(rs.standard_normal(V).astype(np.float32)))
    dG = cu(T(rs.standard_normal(3 * V).astype(np.float32)))
    g0 = cu(T(rs.standard_normal(V).astype(np.float32)))
    vs, w = 0.07, 0.3

    def misaligned(t):   # same values at an address that is 4 (mod 16): every vectorised path is refused
        buf = torch.empty(t.numel() + 1, device=DEV)
        buf[1:] = t
        return buf[1:]

    res = {}
    for name, f in (('vec', lambda t: t.clone()), ('scalar', misaligned)):
        s_, d_, ga, gb = f(sdf), f(dG), f(g0), f(g0)
        G = f(torch.empty(3 * V, device=DEV))
        call('vx_fd_gradient', s_, X, Y, Z, vs, G)
        call('vx_fd_gradient_backward', d_, X, Y, Z, vs, ga)
        call('vx_total_variation_add_grad', s_, ga, None, w, w, w, 1, X, Y, Z, V)
        call('vx_sdf_regularisers_backward', d_, s_, X, Y, Z, vs, w, w, w, gb, None)
        res[name] = (G.clone(), ga.clone(), gb.clone())
    assert torch.equal(res['vec'][1], res['vec'][2]) and torch.equal(res['scalar'][1], res['scalar'][2])
    for a, b in zip(res['vec'], res['scalar']):
        assert torch.equal(a, b)
    # and against the C oracle of the reference TV kernel
    gr = torch.zeros(1, 1, X, Y, Z)
    K.total_variation_add_grad(sdf.cpu().view(1, 1, X, Y, Z), gr, w, w, w, True)
    gt = torch.zeros(1, 1, X, Y, Z, device=DEV)
    tv.total_variation_add_grad(sdf.view(1, 1, X, Y, Z), gt, w, w, w, True)
    close(gt, gr, 1e-6, 1e-7)


@pytest.mark.parametrize('k,sigma', [(5, 0.8), (3, 0.5)])
def test_separable_conv_matches_generic_conv(k, sigma):
    """vx_conv3d_replicate_separable (three 1-D passes) against the generic k^3 kernels, forward and adjoint, on shapes
    smaller than the kernel radius, batches, and with accumulation."""
    from voxurf_b200 import ops
    from voxurf_b200._lib import call
    from voxurf_b200.voxurf_fine import SmoothConv
    sc = SmoothConv(k, sigma)
    for shape in [(2, 1, 12, 10, 14), (1, 1, 2, 3, 2), (1, 3, 1, 1, 5), (1, 1, 33, 17, 40)]:
        rs = np.random.RandomState(k + shape[2])
        x = cu(T(rs.standard_normal(shape).astype(np.float32)))
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        ya = ops.conv3d_replicate(xa, sc.weight_host, k)                                  # generic
        yb = ops.conv3d_replicate(xb, sc.weight_host, k, weight1d=sc.weight1d_host)       # separable
        close(yb, ya, 1e-5, 1e-6)
        go = cu(T(rs.standard_normal(shape).astype(np.float32)))
        ya.backward(go); yb.backward(go)
        close(xb.grad, xa.grad, 1e-5, 2e-6)
        # accumulate form of the adjoint (the fused step adds into the sdf gradient)
        B, X, Y, Z = shape[0] * shape[1], shape[2], shape[3], shape[4]
        acc = cu(T(rs.standard_normal(shape).astype(np.float32)))
        base = acc.clone()
        scratch = torch.empty(2 * x.numel(), device=DEV)
        call('vx_conv3d_replicate_separable', go, B, X, Y, Z, sc.weight1d_host, k, 1, 1, scratch, acc)
        close(acc - base, xa.grad, 1e-5, 2e-6)


@pytest.mark.parametrize('n_rays,n_pts', [(1, 1), (77, 33), (1000, 257), (8192, 64)])
def test_cumdist_thres_bit_exact(n_rays, n_pts):
    """ub360_utils_cuda.cumdist_thres (lib/cuda/ub360_utils_kernel.cu:13-47): warp-per-ray here, thread-per-ray there, same
    float additions in the same order -> identical masks (C oracle; the reference's compiled kernel when present)."""
    from voxurf_b200 import ub360_utils_cuda as ub
    rs = np.random.RandomState(n_rays + n_pts)
    dist = T(np.abs(rs.standard_normal((n_rays, n_pts))).astype(np.float32) * 0.01)
    dist[0, 0] = 10.0
    thres = 0.02
    out = ub.cumdist_thres(dist.cuda(), thres)
    ref = K.cumdist_thres(dist, thres)
    assert out.dtype == torch.bool and torch.equal(out.cpu(), ref)
    assert 0 < int(ref.sum()) < ref.numel() or n_pts == 1
    from oracle import build_ref
    mod = build_ref.load_ref('ub360_utils_cuda')
    if mod is not None:
        assert torch.equal(mod.cumdist_thres(dist.cuda(), thres), out)


def test_adam_blocklive_is_bit_identical_to_dense():
    """vx_adam_step_blocklive (sdf grid): blocks of 128 elements that never saw a gradient are skipped; three steps with
    gradients appearing in new blocks give exactly the dense kernel's parameters and moments."""
    from voxurf_b200._lib import call
    n = 128 * 300
    rs = np.random.RandomState(5)
    p0 = T(rs.standard_normal(n).astype(np.float32)).cuda()
    A = [p0.clone(), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')]
    B = [p0.clone(), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')]
    live = torch.zeros(n // 128, dtype=torch.uint8, device='cuda')
    for step in range(1, 4):
        g = np.zeros(n, np.float32)
        for b in rs.choice(300, 20 * step, replace=False):
            g[b * 128 + rs.randint(0, 128, 5)] = rs.standard_normal(5)
        ga, gb = T(g).cuda(), T(g).cuda()
        bc1, bc2 = 1 - 0.9 ** step, 1 - 0.99 ** step
        args = (0.9, 0.99, 0.1, 0.01, 5e-3 / bc1, float(np.sqrt(bc2)), 1e-8)
        call('vx_adam_step', A[0], ga, A[1], A[2], None, n, *args, 0, 1, None, None, 1, None)
        call('vx_adam_step_blocklive', B[0], gb, B[1], B[2], n, *args, 1, live, None)
        for x, y in zip(A, B):
            assert torch.equal(x, y)
        assert (gb == 0).all()
    assert 0 < int(live.sum()) < 300


# ------------------------------------------------------------------------------------------------ data-parallel exchange kernels
# (one GPU: "ranks" and "peers" are arrays on the same device -- the kernels only see device addresses; the multi-GPU wiring
# over CUDA IPC / NVLink is covered by tests/test_gpu_multi.py on boxes with >= 2 GPUs)
def test_block_nonzero_and_pull_reduce_fixed_order_sum():
    """vx_block_nonzero + vx_pull_reduce (sparse reduce-scatter of the sdf gradient by the slab owner): scale * sum, in rank
    order, of the 128-element blocks the ranks flagged; unflagged blocks keep their content; `out` aliases the own array."""
    from voxurf_b200._lib import call
    rs = np.random.RandomState(11)
    W, nb = 3, 96
    N = nb * 128
    gs, ms = [], []
    for q in range(W):
        g = np.zeros((nb, 128), np.float32)
        for b in rs.choice(nb, 30, replace=False):
            g[b, rs.randint(0, 128, 7)] = rs.standard_normal(7)
        g[5 + q, 3] = 1e-30          # a denormal-ish single element still flags its block
        gs.append(cu(T(g.reshape(-1))))
        m = torch.full((nb,), 7, dtype=torch.uint8, device=DEV)
        call('vx_block_nonzero', gs[q], N, m)
        assert torch.equal(m.bool().cpu(), torch.from_numpy((g != 0).any(1)))
        ms.append(m)
    own = 1
    acc = torch.zeros(N, device=DEV)
    for q in range(W):        # fp32, rank order (adding the exact zeros of unflagged blocks changes nothing)
        acc = acc + gs[q] * ms[q].repeat_interleave(128).float()
    anyb = torch.stack(ms).bool().any(0).repeat_interleave(128)
    want = torch.where(anyb, acc * torch.tensor(np.float32(1.0 / W), device=DEV), gs[own])
    call('vx_pull_reduce', gs[own], N, [g.data_ptr() for g in gs], [m.data_ptr() for m in ms], W, float(np.float32(1.0 / W)))
    assert torch.equal(gs[own], want)
    assert int((~anyb).sum()) > 0


@pytest.mark.parametrize('C', [12, 6])
def test_k0_rows_scatter_equals_per_rank_gather_backward(C):
    """vx_k0_rows_scatter (all ranks' exported k0 rows in one launch, optionally only the corners inside an X-slab) ==
    vx_grid_gather_backward rank by rank (ATen grid_sampler_3d backward arithmetic), then restricted to the slab."""
    from voxurf_b200._lib import call
    rs = np.random.RandomState(3 + C)
    X, Y, Z, W, cap = 10, 12, 8, 3, 400
    V = X * Y * Z
    stride = cap * (3 + C) + 4
    recv = torch.zeros(W, stride, device=DEV)
    counts = [cap, 123, 0]
    for r in range(W):
        recv[r, :cap * 3] = cu(T(rs.uniform(-1.08, 1.08, cap * 3).astype(np.float32)))
        g = rs.standard_normal((cap, C)).astype(np.float32)
        g[::7] = 0                                                   # rows without a gradient are skipped
        recv[r, cap * 3:cap * (3 + C)] = cu(T(g.reshape(-1)))
        recv[r, cap * (3 + C):].view(torch.int32)[0] = counts[r]
    n_words = (V + 31) // 32
    ref, ref_t = torch.zeros(V * C, device=DEV), torch.zeros(n_words, dtype=torch.int32, device=DEV)
    for r in range(W):
        xyz = recv[r, :cap * 3].view(cap, 3).contiguous()
        g = recv[r, cap * 3:cap * (3 + C)].view(cap, C).contiguous()
        n_dev = recv[r, cap * (3 + C):].view(torch.int32).contiguous()
        call('vx_grid_gather_backward', X, Y, Z, C, 1, MN, MX, xyz, None, None, None, None, 0.0, n_dev, cap, g, ref, ref_t)
    assert float(ref.abs().sum()) > 0
    for x_lo, x_hi in ((0, X), (3, 6)):
        got, got_t = torch.zeros(V * C, device=DEV), torch.zeros(n_words, dtype=torch.int32, device=DEV)
        call('vx_k0_rows_scatter', X, Y, Z, C, MN, MX, recv.view(-1), W, cap, x_lo, x_hi, got, got_t)
        want = ref.view(X, -1).clone()
        want[:x_lo] = 0; want[x_hi:] = 0
        close(got.view(X, -1), want, 1e-5, 1e-6)      # (fp32 atomics land in a different order)
        bits = np.unpackbits(ref_t.cpu().numpy().view(np.uint8), bitorder='little')[:V].reshape(X, -1).copy()
        bits[:x_lo] = 0; bits[x_hi:] = 0
        assert (np.unpackbits(got_t.cpu().numpy().view(np.uint8), bitorder='little')[:V].reshape(X, -1) == bits).all()


def test_peer_store_adam_passes_keep_replicas_identical():
    """vx_adam_step_worklist_peers / vx_adam_step_blocklive_peers: the owner's pass on its slice, every updated element also
    stored into the replicas -- bit-identical to the plain passes, replicas == owner afterwards, nothing else touched."""
    from voxurf_b200._lib import call
    rs = np.random.RandomState(21)
    # ---- k0-like grid: C = 12, voxel list from the touched / live bitmaps, the owner steps voxels [lo, hi)
    C, V = 12, 64 * 40
    lo, hi = 64 * 10, 64 * 30
    p0 = cu(T(rs.standard_normal(V * C).astype(np.float32)))
    reps = [p0.clone(), p0.clone()]
    own = [p0.clone(), torch.zeros(V * C, device=DEV), torch.zeros(V * C, device=DEV), torch.zeros(V * C, device=DEV)]
    ref = [t.clone() for t in own]
    work = torch.zeros(V + 1, dtype=torch.int32, device=DEV)
    tch = [torch.zeros(V // 32, dtype=torch.int32, device=DEV) for _ in range(2)]
    liv = [torch.zeros(V // 32, dtype=torch.int32, device=DEV) for _ in range(2)]
    for step in range(1, 4):
        vox = rs.choice(np.arange(lo, hi), 150, replace=False)
        g = np.zeros((V, C), np.float32); g[vox] = rs.standard_normal((150, C))
        bits = np.zeros(V, bool); bits[vox] = True
        words = np.packbits(bits.reshape(-1, 32), axis=1, bitorder='little').view(np.uint32).reshape(-1).astype(np.int64)
        wt = torch.from_numpy(np.where(words >= 2 ** 31, words - 2 ** 32, words)).to(torch.int32)
        bc1, bc2 = 1 - 0.9 ** step, 1 - 0.99 ** step
        args = (0.9, 0.99, 0.1, 0.01, 0.1 / bc1, float(np.sqrt(bc2)), 1e-8)
        for st, t_, l_, peers in ((own, tch[0], liv[0], True), (ref, tch[1], liv[1], False)):
            st[1].copy_(cu(T(g.reshape(-1)))); t_.copy_(wt)
            sl = [x[lo * C:hi * C] for x in st]
            if peers:
                call('vx_adam_step_worklist_peers', *sl, (hi - lo) * C, *args, 1, t_[lo // 32:hi // 32], l_[lo // 32:hi // 32], C, 1, work, None,
                     [r.data_ptr() + 4 * lo * C for r in reps], len(reps))
            else:
                call('vx_adam_step_worklist', *sl, (hi - lo) * C, *args, 1, t_[lo // 32:hi // 32], l_[lo // 32:hi // 32], C, 1, work, None)
        for a, b in zip(own, ref):
            assert torch.equal(a, b)
        for r in reps:
            assert torch.equal(r, own[0])
        assert torch.equal(own[0][:lo * C], p0[:lo * C]) and torch.equal(own[0][hi * C:], p0[hi * C:])
    # ---- sdf-like grid: block-live pass on the slab [lo, hi) of a single-channel array
    n, lo, hi = 128 * 200, 128 * 50, 128 * 150
    p0 = cu(T(rs.standard_normal(n).astype(np.float32)))
    reps = [p0.clone(), p0.clone(), p0.clone()]
    own = [p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)]
    ref = [t.clone() for t in own]
    live = [torch.zeros(n // 128, dtype=torch.uint8, device=DEV) for _ in range(2)]
    for step in range(1, 4):
        g = np.zeros(n, np.float32)
        for b in rs.choice(np.arange(50, 150), 15, replace=False):
            g[b * 128 + rs.randint(0, 128, 4)] = rs.standard_normal(4)
        bc1, bc2 = 1 - 0.9 ** step, 1 - 0.99 ** step
        args = (0.9, 0.99, 0.1, 0.01, 5e-3 / bc1, float(np.sqrt(bc2)), 1e-8)
        own[1].copy_(cu(T(g))); ref[1].copy_(cu(T(g)))
        call('vx_adam_step_blocklive_peers', *[x[lo:hi] for x in own], hi - lo, *args, 1, live[0][lo // 128:hi // 128], None,
             [r.data_ptr() + 4 * lo for r in reps], len(reps))
        call('vx_adam_step_blocklive', *[x[lo:hi] for x in ref], hi - lo, *args, 1, live[1][lo // 128:hi // 128], None)
        for a, b in zip(own, ref):
            assert torch.equal(a, b)
        for r in reps:
            assert torch.equal(r, own[0])
    assert 0 < int(live[0].sum()) < 100
