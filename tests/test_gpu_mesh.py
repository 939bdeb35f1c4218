"""GPU: the mesh export after the field query -- vertex colouring against the reference's own
mesh_color_forward (tests/golden/r2_extras.npz) and the oracle; GPU marching cubes against the table-free CPU checks of
oracle/marching_ref.py and its pure-Python replay of the extraction."""
import numpy as np
import pytest
import torch

from oracle import marching_ref as MR
from oracle import voxurf_ref as R
from voxurf_b200 import synthetic as S
from tests.helpers import T, load_golden, oracle_fine_model, product_fine_model

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def test_vertex_colours_match_reference():
    g = load_golden('r2_extras.npz')
    sc = S.make_fine_scene(24, 6, 32, seed=5, mask_G=12)
    for cl in (False, True):
        m = product_fine_model(sc, k0_channels_last=cl)
        rgb = m.mesh_color_forward(T(g['mc_pts']).to(DEV))
        np.testing.assert_allclose(rgb.cpu().numpy(), g['mc_rgb'], rtol=1e-5, atol=3e-6)
    # more points than one row chunk, against the oracle
    om = oracle_fine_model(sc, requires_grad=False)
    rs = np.random.RandomState(3)
    d = rs.standard_normal((70000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = T((d * (0.5 + 0.05 * rs.standard_normal((70000, 1)))).astype(np.float32))
    with torch.no_grad():
        ref = R.mesh_color_forward(om, pts)
    rgb = m.mesh_color_forward(pts.to(DEV))
    np.testing.assert_allclose(rgb.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=3e-6)


@pytest.mark.parametrize('kind,n', [('sphere', 20), ('noise', 12), ('bumpy', 33)])
def test_marching_cubes_matches_cpu_replay_and_is_watertight(kind, n):
    from voxurf_b200 import marching as M
    rs = np.random.RandomState(n)
    g = np.stack(np.meshgrid(*[np.linspace(-1, 1, n)] * 3, indexing='ij'), -1)
    if kind == 'sphere':
        u = 0.6 - np.linalg.norm(g, axis=-1)
    elif kind == 'noise':
        u = rs.standard_normal((n, n, n))
    else:
        u = 0.55 - np.linalg.norm(g + 0.03, axis=-1) + 0.2 * np.sin(5 * g[..., 0]) * np.sin(4 * g[..., 1] + 1) * np.cos(3 * g[..., 2])
    u = u.astype(np.float32)
    u[0] = u[-1] = -1; u[:, 0] = u[:, -1] = -1; u[:, :, 0] = u[:, :, -1] = -1       # level set away from the lattice boundary
    thr = 0.0 if kind != 'sphere' else 0.05
    v, t = M.marching_cubes(T(u).to(DEV), thr)
    v, t = v.cpu().numpy(), t.cpu().numpy()
    MR.check_mesh(u, thr, v, t)                                           # table-free: vertex set, closedness, orientation
    tab, cnt = M.build_tables()
    v_ref, t_ref = MR.mc_numpy(u, thr, tab, M._CORNER, M._EDGES)          # pure-Python replay: same order, same numbers
    assert np.array_equal(v, v_ref) and np.array_equal(t, t_ref)
    p = v[t]
    vol = np.einsum('ij,ij->i', p[:, 0], np.cross(p[:, 1], p[:, 2])).sum() / 6
    assert vol > 0                                                        # normals point from u > thr to u < thr
    if kind == 'sphere':
        r_idx = (0.6 - thr) * (n - 1) / 2
        assert abs(vol - 4 / 3 * np.pi * r_idx ** 3) < 0.03 * vol and len(v) - 3 * len(t) // 2 + len(t) == 2


def test_extract_geometry_end_to_end():
    """field query -> marching cubes -> world coordinates -> vertex colours, all on the GPU (run.py:873-909)."""
    sc = S.make_fine_scene(32, 6, 32, seed=8, mask_G=16)
    m = product_fine_model(sc, k0_channels_last=True)
    verts, tris = m.extract_geometry(resolution=48, threshold=0.0)
    assert verts.shape[1] == 3 and tris.shape[1] == 3 and len(tris) > 100
    r = np.linalg.norm(verts, axis=1)
    assert abs(np.median(r) - 0.5) < 0.03            # the scene is a (noisy) sphere of radius 0.5
    u = m.query_sdf_field(48).cpu().numpy()
    MR.check_mesh(u, 0.0, (verts + 1.0) / 2.0 * 47.0, tris)
    rgb = m.mesh_color_forward(torch.from_numpy(verts).to(DEV))
    assert rgb.shape == (len(verts), 3) and torch.isfinite(rgb).all() and float(rgb.min()) >= 0 and float(rgb.max()) <= 1
