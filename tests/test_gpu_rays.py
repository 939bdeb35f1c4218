"""GPU: the caller side of the path -- view rays, hit_coarse_geo and the training-ray gathering of
voxurf_b200/rays.py against the oracle and against vectors produced by the reference's own functions
(tests/golden/rays.npz).  Float outputs: rtol 2e-6 (one rounding per op, same order); hit flags, per-view counts and
the order of the kept rows: exact."""
import numpy as np
import pytest
import torch

from oracle import voxurf_ref as R
from voxurf_b200 import synthetic as S
from tests.helpers import T, load_golden, oracle_fine_model, product_fine_model

pytestmark = pytest.mark.gpu
DEV = 'cuda'
XYZ_MIN, XYZ_MAX = torch.tensor([-1., -1., -1.]), torch.tensor([1., 1., 1.])


def close(a, b, rtol=2e-6, atol=1e-6):
    np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy() if torch.is_tensor(b) else b, rtol=rtol, atol=atol)


@pytest.mark.parametrize('name', list(S.RAY_CASES))
def test_get_rays_of_a_view_matches_reference_and_oracle(name):
    from voxurf_b200 import voxurf_fine as vf
    kw = S.RAY_CASES[name]
    g = load_golden('rays.npz')
    H, W, K, c2w = S.make_view(seed=kw['seed'], H=kw['H'], W=kw['W'], inverse_y=kw['inverse_y'])
    ro, rd, vd = vf.get_rays_of_a_view(H, W, K, T(c2w).to(DEV), kw['ndc'], kw['inverse_y'], kw['flip_x'], kw['flip_y'], mode=kw['mode'])
    assert ro.shape == (H, W, 3) and rd.shape == (H, W, 3) and vd.shape == (H, W, 3)
    close(ro, g[name + '_rays_o']); close(rd, g[name + '_rays_d']); close(vd, g[name + '_viewdirs'], 2e-6, 1e-7)
    oro, ord_, ovd = R.view_rays(H, W, T(K), T(c2w), ndc=kw['ndc'], inverse_y=kw['inverse_y'], flip_x=kw['flip_x'],
                                 flip_y=kw['flip_y'], mode=kw['mode'])
    close(ro, oro); close(rd, ord_); close(vd, ovd, 2e-6, 1e-7)
    if not kw['ndc']:
        o2, d2 = vf.get_rays(H, W, K, T(c2w).to(DEV), kw['inverse_y'], kw['flip_x'], kw['flip_y'], mode=kw['mode'])
        assert torch.equal(o2, ro) and torch.equal(d2, rd)


def test_get_rays_random_mode_and_bad_mode():
    from voxurf_b200 import voxurf_fine as vf
    H, W, K, c2w = S.make_view(seed=9, H=11, W=13)
    torch.manual_seed(5)
    ro, rd, vd = vf.get_rays_of_a_view(H, W, K, T(c2w).to(DEV), False, False, True, False, mode='random')
    torch.manual_seed(5)
    ji = torch.rand(H, W, device=DEV).cpu(); jj = torch.rand(H, W, device=DEV).cpu()
    oro, ord_, ovd = R.view_rays(H, W, T(K), T(c2w), flip_x=True, mode='random', jitter=(ji, jj))
    close(rd, ord_); close(vd, ovd, 2e-6, 1e-7)
    with pytest.raises(NotImplementedError):
        vf.get_rays(H, W, K, T(c2w).to(DEV), False, False, False, mode='nope')


def test_hit_coarse_geo_exact():
    sc = S.make_fine_scene(20, 6, 32, seed=3, mask_G=12)
    m = product_fine_model(sc)
    om = oracle_fine_model(sc, requires_grad=False)
    for seed, (h, w) in enumerate([(40, 48), (17, 5)]):
        H, W, K, c2w = S.make_view(seed=60 + seed, H=h, W=w)
        ro, rd, _ = R.view_rays(H, W, T(K), T(c2w))
        ro = ro.contiguous()
        # far = 0.5 would end every ray before the box if it were honoured: the reference overrides it with 1e9 (hazard 4)
        hit = m.hit_coarse_geo(ro.to(DEV), rd.to(DEV), near=0.3, far=0.5, stepsize=0.5, bg=0.0)
        ref = R.hit_coarse_geo(om['mask_cache'], ro, rd, XYZ_MIN, XYZ_MAX, 0.3, 0.5, om['voxel_size'])
        assert hit.shape == (H, W) and hit.dtype == torch.bool
        assert torch.equal(hit.cpu(), ref) and 0 < int(ref.sum()) < ref.numel()
    assert m.hit_coarse_geo(torch.zeros(0, 3, device=DEV), torch.zeros(0, 3, device=DEV), near=0.3, far=6.0, stepsize=0.5).shape == (0,)


def test_get_training_rays_in_maskcache_sampling_matches_reference():
    from voxurf_b200 import voxurf_fine as vf
    g = load_golden('rays.npz')
    sc = S.make_fine_scene(20, 6, 32, seed=3, mask_G=12)
    m = product_fine_model(sc)
    views = [S.make_view(seed=40 + i, H=h, W=w, inverse_y=False) for i, (h, w) in enumerate(S.TRAIN_VIEW_SIZES)]
    imgs = [T(S.make_image(h, w, seed=i)).to(DEV) for i, (h, w, _, _) in enumerate(views)]
    poses = torch.stack([T(v[3]) for v in views]).to(DEV)
    HW = np.array([(v[0], v[1]) for v in views]); Ks = np.stack([v[2] for v in views])
    rk = dict(near=0.3, far=6.0, bg=0.0, stepsize=0.5)
    rgb, ro, rd, vd, imsz = vf.get_training_rays_in_maskcache_sampling(imgs, poses, HW, Ks, False, False, False, False, m, rk)
    assert [int(n) for n in imsz] == g['tr_imsz'].tolist()
    assert np.array_equal(rgb.cpu().numpy(), g['tr_rgb'])       # kept pixels in order == kept rays in order
    close(ro, g['tr_rays_o']); close(rd, g['tr_rays_d']); close(vd, g['tr_viewdirs'], 2e-6, 1e-7)
    # the unfiltered gatherers
    rgb2, ro2, rd2, vd2, imsz2 = vf.get_training_rays_flatten(imgs, poses, HW, Ks, False, False, False, False)
    assert imsz2 == [h * w for h, w in S.TRAIN_VIEW_SIZES] and ro2.shape == (sum(imsz2), 3)
    top = 0
    for (H, W, K, c2w), n in zip(views, imsz2):
        oro, ord_, ovd = R.view_rays(H, W, T(K), T(c2w))
        close(rd2[top:top + n], ord_.reshape(-1, 3)); close(vd2[top:top + n], ovd.reshape(-1, 3), 2e-6, 1e-7)
        top += n
    same = [0, 2]
    rgb_s = torch.stack([imgs[i] for i in same])
    out = vf.get_training_rays(rgb_s, poses[same], HW[same], Ks[same], False, False, False, False)
    assert out[1].shape == (2, 40, 48, 3) and out[4] == [1, 1]
    close(out[2][1], R.view_rays(40, 48, T(views[2][2]), T(views[2][3]))[1])
    it = vf.batch_indices_generator(10, 4)
    assert next(it).shape == (4,) and next(it).dtype == torch.int64
