"""Round-2 golden vectors, again produced by the REFERENCE'S OWN Python code on CPU (same stubbing as make_golden.py):

    python tests/golden/make_golden_r2.py        (build container only: needs /root/reference)

  r2_extras.npz
    ctv_*    lib/voxurf_coarse.py:300-320,702-715  coarse `total_variation` (sum / 3 / mask.sum() -- NOT the fine file's
             per-axis means), density_total_variation(sdf_tv, smooth_grad_tv), k0_total_variation + their gradients
    sv_*     lib/voxurf_fine.py:384-397 + lib/grid.py:60-65  scale_volume_grid 20^3 -> 28^3: sdf, k0, nonempty mask
    field_*  lib/voxurf_fine.py:894-910 + lib/dvgo_ori.py:679-693  extract_fields of -smooth_k3(sdf) on a 24^3 lattice,
             and the trilinear gradient field of the same smoothed grid (grid_sampler(sample_grad=True)) on that lattice
    mc_*     lib/voxurf_fine.py:804-892  mesh_color_forward on points near the surface (vertex colouring)
    schema_* sorted "key:shape" of model.state_dict() for a fine (smooth_ksize=5) and a coarse model with a mask cache
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402
from voxurf_b200 import synthetic as S  # noqa: E402

T = MG.T


def main():
    MG._install_stubs()
    sys.path.insert(0, '/root/reference')
    import lib.voxurf_fine as vf
    import lib.voxurf_coarse as vc
    import lib.dvgo_ori as dvgo
    torch.manual_seed(0)
    tmp = '/tmp/voxurf_golden'
    os.makedirs(tmp, exist_ok=True)
    out = MG.Snap()

    # ---------------------------------------------------------------- coarse regularisers (autograd forms)
    Gc, Cc, Wc = 16, 12, 32
    scc = S.make_coarse_scene(Gc, Cc, Wc, seed=4, mask_G=12)
    ckpt_c = os.path.join(tmp, 'mask_coarse.tar')
    MG.write_mask_ckpt(scc, ckpt_c)
    mc = MG.build_coarse(vc, scc, ckpt_c)
    mc._set_nonempty_mask()
    out['ctv_value'] = vc.total_variation(mc.sdf.grid, mc.nonempty_mask)
    mc.gradient = mc.neus_sdf_gradient(sdf=mc.sdf.grid)
    mc.sdf.grid.grad = None
    tvl = mc.density_total_variation(sdf_tv=0.1, smooth_grad_tv=0.2)
    tvl.backward()
    out['ctv_density_value'] = tvl.detach()
    out['ctv_density_grad'] = mc.sdf.grid.grad.clone()
    mc.k0.grid.grad = None
    k0l = mc.k0_total_variation()
    k0l.backward()
    out['ctv_k0_value'] = k0l.detach()
    out['ctv_k0_grad'] = mc.k0.grid.grad.clone()

    # ---------------------------------------------------------------- fine: progressive growing
    G, C, W = 20, 6, 32
    sc = S.make_fine_scene(G, C, W, seed=3, mask_G=12)
    ckpt = os.path.join(tmp, 'mask_fine.tar')
    MG.write_mask_ckpt(sc, ckpt)
    m = MG.build_fine(vf, sc, ckpt)
    m._set_nonempty_mask()
    m.scale_volume_grid(23000)   # 28.4^3: clear of the float32 pow/floor edge at exact cubes
    out.update(sv_world_size=m.world_size, sv_sdf=m.sdf.grid.data, sv_k0=m.k0.grid.data, sv_nonempty=m.nonempty_mask,
               sv_voxel_size=m.voxel_size)

    # ---------------------------------------------------------------- fine: mesh field + gradient field + vertex colours
    sc = S.make_fine_scene(24, 6, 32, seed=5, mask_G=12)
    MG.write_mask_ckpt(sc, ckpt)
    m = MG.build_fine(vf, sc, ckpt)
    m._set_nonempty_mask()
    res = 24
    with torch.no_grad():
        m.init_smooth_conv_test_k3(sigma=0.5)
        sdf_grid = m.smooth_conv_test_k3(m.sdf.grid)
        u = dvgo.extract_fields(m.xyz_min, m.xyz_max, res, lambda pts: m.grid_sampler(pts, -sdf_grid), N=10)
        out['field_u'] = torch.from_numpy(u)
        u_raw = dvgo.extract_fields(m.xyz_min, m.xyz_max, res, lambda pts: m.grid_sampler(pts, -m.sdf.grid), N=64)
        out['field_u_raw'] = torch.from_numpy(u_raw)
        ax = [torch.linspace(float(m.xyz_min[i]), float(m.xyz_max[i]), res) for i in range(3)]
        xx, yy, zz = torch.meshgrid(*ax, indexing='ij')
        pts = torch.stack([xx, yy, zz], -1).reshape(-1, 3)
        s, g, _ = m.grid_sampler(pts, sdf_grid, sample_ret=True, sample_grad=True, displace=1.0)
        out['field_sdf'] = s.reshape(res, res, res)
        out['field_grad'] = g.reshape(res, res, res, 3)
        # vertex colouring: points on / near the sphere surface r = 0.5 (run.py:873-909 feeds mesh vertices)
        rs = np.random.RandomState(9)
        d = rs.standard_normal((200, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        vpts = T((d * (0.5 + 0.03 * rs.standard_normal((200, 1)))).astype(np.float32))
        out['mc_pts'] = vpts
        out['mc_rgb'] = m.mesh_color_forward(vpts)
    # ---------------------------------------------------------------- state_dict schemas (checkpoint compatibility)
    def schema(model):
        return np.array(sorted(f'{k}:{tuple(v.shape)}' for k, v in model.state_dict().items()))
    sc = S.make_fine_scene(20, 6, 32, seed=3, mask_G=12)
    MG.write_mask_ckpt(sc, ckpt)
    cfg = {k: v for k, v in S.FINE_CFG.items() if k not in ('stepsize',)}
    mf = vf.Voxurf(xyz_min=[-1., -1., -1.], xyz_max=[1., 1., 1.], num_voxels=20 ** 3, num_voxels_base=20 ** 3,
                   mask_cache_path=ckpt, rgbnet_dim=6, rgbnet_width=32, smooth_ksize=5, smooth_sigma=0.8, **cfg)
    out['schema_fine'] = schema(mf)
    out['schema_coarse'] = schema(mc)
    np.savez_compressed(os.path.join(HERE, 'r2_extras.npz'), **MG.npy(out))
    print('r2_extras.npz', os.path.getsize(os.path.join(HERE, 'r2_extras.npz')))


if __name__ == '__main__':
    main()
