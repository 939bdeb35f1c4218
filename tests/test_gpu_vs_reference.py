"""GPU: this repo's kernels against the REFERENCE'S OWN CUDA kernels (oracle/_ref: the unmodified sources of
/root/reference/lib/cuda compiled by oracle/build_ref.py), and the C oracle against the same -- this is what
pins oracle/ref_kernels.c.  Integer / index / bool outputs must be bit-exact."""
import numpy as np
import pytest
import torch

from oracle import kernels as K
from oracle.build_ref import load_ref
from voxurf_b200 import synthetic as S
from tests.helpers import T

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _ref(name):
    mod = load_ref(name)
    if mod is None:
        pytest.skip(f'oracle/_ref/{name} not built (needs /root/reference at build time)')
    return mod


def exact(a, b, msg=''):
    a, b = a.detach().cpu(), b.detach().cpu()
    assert a.shape == b.shape and a.dtype == b.dtype, (msg, a.shape, b.shape, a.dtype, b.dtype)
    assert torch.equal(a, b), (msg, (a != b).sum().item())


def close(a, b, rtol, atol, msg=''):
    np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy(), rtol=rtol, atol=atol, err_msg=msg)


@pytest.mark.parametrize('n,G', [(8192, 256), (8192, 96), (513, 160)])
def test_sampling_three_way_bit_exact(n, G):
    ref = _ref('render_utils_cuda')
    from voxurf_b200 import render_utils_cuda as ru
    o, d, _ = (T(x) for x in S.make_rays(n, seed=G))
    d[0, 1] = 0.0
    o[2] = torch.tensor([5.0, 5.0, 5.0])
    mn, mx = torch.tensor([-1., -1., -1.]), torch.tensor([1., 1., 1.])
    stepdist = float(np.float32(0.5 * 2.0 / G))
    a = ref.sample_pts_on_rays(o.to(DEV), d.to(DEV), mn.to(DEV), mx.to(DEV), 0.3, 1e9, stepdist)
    b = ru.sample_pts_on_rays(o.to(DEV), d.to(DEV), mn.to(DEV), mx.to(DEV), 0.3, 1e9, stepdist)
    c = K.sample_pts_on_rays(o, d, mn, mx, 0.3, 1e9, stepdist)
    for x, y, z, name in zip(a, b, c, ['pts', 'mask_outbbox', 'ray_id', 'step_id', 'N_steps', 't_min', 't_max']):
        exact(y, x, 'ours vs reference: ' + name)
        exact(z, x, 'C oracle vs reference: ' + name)


def test_maskcache_lookup_three_way():
    ref = _ref('render_utils_cuda')
    from voxurf_b200 import render_utils_cuda as ru
    from oracle import voxurf_ref as R
    rs = np.random.RandomState(3)
    world = T(rs.uniform(0, 1, (33, 40, 37)) > 0.5)
    q = T(rs.uniform(-1.2, 1.2, (200000, 3)).astype(np.float32))
    scale, shift = R.mask_grid_params(world.shape, torch.tensor([-1., -1., -1.]), torch.tensor([1., 1., 1.]))
    a = ref.maskcache_lookup(world.to(DEV), q.to(DEV), scale.to(DEV), shift.to(DEV))
    exact(ru.maskcache_lookup(world.to(DEV), q.to(DEV), scale.to(DEV), shift.to(DEV)), a)
    exact(K.maskcache_lookup(world, q, scale, shift), a)


@pytest.mark.parametrize('n_rays', [8192, 100])
def test_alpha2weight_three_way_bit_exact(n_rays):
    ref = _ref('render_utils_cuda')
    from voxurf_b200 import render_utils_cuda as ru
    rs = np.random.RandomState(n_rays)
    lens = rs.randint(0, 400, n_rays); lens[rs.uniform(0, 1, n_rays) < 0.1] = 0
    rid = torch.from_numpy(np.repeat(np.arange(n_rays), lens).astype(np.int64))
    alpha = T((rs.uniform(0, 1, rid.shape[0]) ** 4 * 0.5).astype(np.float32))
    a = ref.alpha2weight(alpha.to(DEV), rid.to(DEV), n_rays)
    b = ru.alpha2weight(alpha.to(DEV), rid.to(DEV), n_rays)
    c = K.alpha2weight(alpha, rid, n_rays)
    for x, y, z, name in zip(a, b, c, ['weight', 'T', 'alphainv_last', 'i_start', 'i_end']):
        exact(y, x, 'ours vs reference: ' + name)
        exact(z, x, 'C oracle vs reference: ' + name)
    gw = T(rs.standard_normal(rid.shape[0]).astype(np.float32)).to(DEV)
    gl = T(rs.standard_normal(n_rays).astype(np.float32)).to(DEV)
    ga = ref.alpha2weight_backward(alpha.to(DEV), *a, n_rays, gw, gl)
    exact(ru.alpha2weight_backward(alpha.to(DEV), *b, n_rays, gw, gl), ga, 'ours vs reference: grad')
    exact(K.alpha2weight_backward(alpha, *c, n_rays, gw.cpu(), gl.cpu()), ga, 'C oracle vs reference: grad')


def test_raw2alpha_vs_reference():
    ref = _ref('render_utils_cuda')
    from voxurf_b200 import render_utils_cuda as ru
    d = (torch.randn(100000) * 6).to(DEV)
    a, b = ref.raw2alpha(d, -4.0, 0.5), ru.raw2alpha(d, -4.0, 0.5)
    exact(b[0], a[0]); exact(b[1], a[1])
    gb = torch.randn(100000, device=DEV)
    exact(ru.raw2alpha_backward(b[0], gb, 0.5), ref.raw2alpha_backward(a[0], gb, 0.5))
    close(K.raw2alpha(d.cpu(), -4.0, 0.5)[1], a[1], 1e-5, 1e-7)


@pytest.mark.parametrize('dense', [True, False])
@pytest.mark.parametrize('shape', [(1, 2, 40, 33, 47), (1, 1, 40, 33, 48)])   # the second takes the z-vectorised kernel
def test_total_variation_vs_reference(dense, shape):
    ref = _ref('total_variation_cuda')
    from voxurf_b200 import total_variation_cuda as tv
    rs = np.random.RandomState(5)
    p = T(rs.standard_normal(shape).astype(np.float32) * 2)
    g = T(rs.standard_normal(p.shape).astype(np.float32)); g[rs.uniform(0, 1, g.shape) < 0.5] = 0
    mk = T((rs.uniform(0, 1, p.shape) > 0.3).astype(np.float32))
    a, b, c = g.clone().to(DEV), g.clone().to(DEV), g.clone()
    ref.total_variation_add_grad(p.to(DEV), a, 0.3, 0.5, 0.7, dense)
    tv.total_variation_add_grad(p.to(DEV), b, 0.3, 0.5, 0.7, dense)
    K.total_variation_add_grad(p, c, 0.3, 0.5, 0.7, dense)
    close(b, a, 1e-6, 1e-7, 'ours vs reference'); close(c, a, 1e-6, 1e-7, 'C oracle vs reference')
    a, b, c = g.clone().to(DEV), g.clone().to(DEV), g.clone()
    ref.total_variation_add_grad_new(p.to(DEV), a, mk.to(DEV), 0.3, 0.5, 0.7, dense)
    tv.total_variation_add_grad_new(p.to(DEV), b, mk.to(DEV), 0.3, 0.5, 0.7, dense)
    K.total_variation_add_grad(p, c, 0.3, 0.5, 0.7, dense, mask=mk)
    close(b, a, 1e-6, 1e-7, 'ours vs reference (masked)'); close(c, a, 1e-6, 1e-7, 'C oracle vs reference (masked)')


@pytest.mark.parametrize('mode', [0, 1, 2])
def test_adam_upd_vs_reference(mode):
    ref = _ref('adam_upd_cuda')
    from voxurf_b200 import adam_upd_cuda as ad
    rs = np.random.RandomState(7)
    n = 1 << 20
    p, m, v = (T(rs.standard_normal(n).astype(np.float32)) for _ in range(3))
    v = v.abs()
    perlr = T(rs.uniform(0, 1, n).astype(np.float32)).to(DEV)
    A = [x.clone().to(DEV) for x in (p, m, v)]
    B = [x.clone().to(DEV) for x in (p, m, v)]
    C = [x.clone() for x in (p, m, v)]
    for step in (1, 7):
        g = T(rs.standard_normal(n).astype(np.float32)); g[::3] = 0
        gd = g.to(DEV)
        if mode == 0:
            ref.adam_upd(A[0], gd, A[1], A[2], step, 0.9, 0.99, 0.1, 1e-8); ad.adam_upd(B[0], gd, B[1], B[2], step, 0.9, 0.99, 0.1, 1e-8)
        elif mode == 1:
            ref.masked_adam_upd(A[0], gd, A[1], A[2], step, 0.9, 0.99, 0.1, 1e-8); ad.masked_adam_upd(B[0], gd, B[1], B[2], step, 0.9, 0.99, 0.1, 1e-8)
        else:
            ref.adam_upd_with_perlr(A[0], gd, A[1], A[2], perlr, step, 0.9, 0.99, 0.1, 1e-8)
            ad.adam_upd_with_perlr(B[0], gd, B[1], B[2], perlr, step, 0.9, 0.99, 0.1, 1e-8)
        K.adam_upd(C[0], g, C[1], C[2], step, 0.9, 0.99, 0.1, 1e-8, mode=mode, perlr=perlr.cpu() if mode == 2 else None)
        for x, y, z in zip(A, B, C):
            close(y, x, 1e-6, 1e-6, 'ours vs reference'); close(z, x, 1e-6, 1e-6, 'C oracle vs reference')
