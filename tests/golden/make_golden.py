"""Generates tests/golden/*.npz by running the REFERENCE'S OWN Python code on CPU.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py

How: /root/reference is put on sys.path and `lib.voxurf_fine`, `lib.voxurf_coarse`, `lib.grid`,
`lib.utils`, `lib.dvgo_ori` are imported unmodified.  Their missing third-party imports
(matplotlib, torch_scatter, mcubes, plyfile, imageio, skimage, trimesh) are replaced by stubs,
`torch.utils.cpp_extension.load` returns a module backed by oracle/kernels.py (the C restatement of
the reference's CUDA operators -- the only non-reference arithmetic in these vectors, pinned
separately on the GPU against oracle/_ref), and `.cuda()` is a no-op.  Inputs come from the seeded
recipes in voxurf_b200/synthetic.py, so only outputs are stored.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import kernels as OK  # noqa: E402
from voxurf_b200 import synthetic as S  # noqa: E402


def _install_stubs():
    class _Anything(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith('__'):
                raise AttributeError(name)
            return _Anything(self.__name__ + '.' + name)

        def __call__(self, *a, **k):
            return None

    for name in ['matplotlib', 'matplotlib.pyplot', 'matplotlib.cm', 'mcubes', 'plyfile', 'imageio', 'skimage',
                 'skimage.measure', 'trimesh', 'cv2', 'scipy.signal']:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = _Anything(name)
    ts = types.ModuleType('torch_scatter')

    def segment_coo(src, index, out=None, reduce='sum'):
        assert reduce == 'sum'
        return out.index_add_(0, index, src)
    ts.segment_coo = segment_coo
    sys.modules['torch_scatter'] = ts

    ru = types.ModuleType('render_utils_cuda')
    ru.sample_pts_on_rays = lambda *a: list(OK.sample_pts_on_rays(*a))
    ru.infer_t_minmax = lambda *a: list(OK.infer_t_minmax(*a))
    ru.infer_n_samples = OK.infer_n_samples
    ru.infer_ray_start_dir = lambda *a: list(OK.infer_ray_start_dir(*a))
    ru.maskcache_lookup = OK.maskcache_lookup
    ru.alpha2weight = lambda *a: list(OK.alpha2weight(*a))
    ru.alpha2weight_backward = OK.alpha2weight_backward
    tv = types.ModuleType('total_variation_cuda')
    tv.total_variation_add_grad = lambda p, g, wx, wy, wz, dense: OK.total_variation_add_grad(p.data, g, wx, wy, wz, dense)
    tv.total_variation_add_grad_new = lambda p, g, m, wx, wy, wz, dense: OK.total_variation_add_grad(p.data, g, wx, wy, wz, dense, mask=m)
    import torch.utils.cpp_extension as ce
    ce.load = lambda name, **kw: {'render_utils_cuda': ru, 'total_variation_cuda': tv}[name]
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x))


class Snap(dict):
    """dict that clones tensors on insertion (later in-place updates must not leak into the vectors)."""

    def __setitem__(self, k, v):
        super().__setitem__(k, v.detach().clone() if torch.is_tensor(v) else v)

    def __init__(self, *a, **kw):
        super().__init__()
        self.update(*a, **kw)

    def update(self, *a, **kw):
        for k, v in dict(*a, **kw).items():
            self[k] = v


def npy(d):
    out = {}
    for k, v in d.items():
        if v is None:
            continue
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    return out


def set_mlp(seq, layers):
    lin = [m for m in seq.modules() if isinstance(m, torch.nn.Linear)]
    assert len(lin) == len(layers)
    for m, (W, b) in zip(lin, layers):
        m.weight.data = T(W).clone()
        m.bias.data = T(b).clone()


def write_mask_ckpt(sc, path, xyz_min=(-1., -1., -1.), xyz_max=(1., 1., 1.)):
    torch.save({'MaskCache_kwargs': {'xyz_min': list(xyz_min), 'xyz_max': list(xyz_max),
                                     'act_shift': sc['mask_act_shift'],
                                     'voxel_size_ratio': sc['mask_voxel_size_ratio'], 'nearest': False},
                'model_state_dict': {'density': T(sc['mask_density'])}}, path)


def build_fine(vf, sc, ckpt):
    m = vf.Voxurf(xyz_min=[-1., -1., -1.], xyz_max=[1., 1., 1.], num_voxels=sc['G'] ** 3, num_voxels_base=sc['G'] ** 3,
                  mask_cache_path=ckpt, rgbnet_dim=sc['C'], rgbnet_width=sc['width'],
                  **{k: v for k, v in S.FINE_CFG.items() if k not in ('stepsize',)})
    assert tuple(m.world_size.tolist()) == (sc['G'],) * 3, m.world_size
    m.sdf.grid.data = T(sc['sdf']).clone()
    m.k0.grid.data = T(sc['k0']).clone()
    set_mlp(m.rgbnet, sc['rgbnet'])
    set_mlp(m.k_rgbnet, sc['k_rgbnet'])
    return m


def build_coarse(vc, sc, ckpt):
    cfg = dict(S.COARSE_CFG)
    cfg.pop('stepsize')
    cfg['rgbnet_dim'] = sc['C']
    cfg['rgbnet_width'] = sc['width']
    m = vc.Voxurf(xyz_min=[-1., -1., -1.], xyz_max=[1., 1., 1.], num_voxels=sc['G'] ** 3, num_voxels_base=sc['G'] ** 3,
                  mask_cache_path=ckpt, rgbnet_direct=True, **cfg)
    assert tuple(m.world_size.tolist()) == (sc['G'],) * 3, m.world_size
    m.sdf.grid.data = T(sc['sdf']).clone()
    m.k0.grid.data = T(sc['k0']).clone()
    set_mlp(m.rgbnet, sc['rgbnet'])
    return m


def main():
    _install_stubs()
    sys.path.insert(0, '/root/reference')
    import lib.voxurf_fine as vf
    import lib.voxurf_coarse as vc
    import lib.grid as rgrid
    import lib.utils as rutils
    import lib.dvgo_ori as dvgo
    torch.manual_seed(0)
    tmp = '/tmp/voxurf_golden'
    os.makedirs(tmp, exist_ok=True)

    # ---------------------------------------------------------------- fine model: per-op vectors
    G, C, W = 20, 6, 32
    sc = S.make_fine_scene(G, C, W, seed=3, mask_G=12)
    ckpt = os.path.join(tmp, 'mask_fine.tar')
    write_mask_ckpt(sc, ckpt)
    m = build_fine(vf, sc, ckpt)
    m.sdf.grid.data = T(sc['sdf']).clone()   # undo _set_nonempty_mask's overwrite for the op-level vectors
    rs = np.random.RandomState(11)
    pts = T(rs.uniform(-1.0, 1.0, (300, 3)).astype(np.float32))
    pts[:8] = T(np.array([[-1, -1, -1], [1, 1, 1], [1, -1, 0.3], [0.999, 0.2, -0.999], [0, 0, 0],
                          [-0.95, 0.95, 0.5], [0.5, 1.0, -1.0], [-1.0, 0.1, 0.2]], np.float32))
    ops = Snap()
    with torch.no_grad():
        sdf, grad, feat = m.grid_sampler(pts, m.sdf.grid, sample_ret=True, sample_grad=True, displace=1.0)
        ops.update(gs_sdf=sdf, gs_grad=grad, gs_feat=feat)
        f4, g4 = m.sample_sdfs(pts, m.sdf.grid, displace_list=[0.5, 1.0, 1.5, 2.0], use_grad_norm=True)
        ops.update(ss_feat=f4, ss_grad=g4)
        f4r, g4r = m.sample_sdfs(pts, m.sdf.grid, displace_list=[0.5, 1.0, 1.5, 2.0], use_grad_norm=False)
        ops.update(ss_grad_raw=g4r)
        ops['k0'] = m.k0(pts)
        ops['fd_gradient'] = m.neus_sdf_gradient()
        ops['smooth_k5'] = m._gaussian_3dconv(5, 0.8)(m.sdf.grid)
        ops['smooth_k3'] = m._gaussian_3dconv(3, 0.5)(m.sdf.grid)
        ops['smooth_k5_weight'] = m._gaussian_3dconv(5, 0.8).weight
        ops['tv_smooth_weight'] = m.tv_smooth_conv.weight
        ops['mask_cache'] = m.mask_cache(pts)
        ops['mask_cache_density'] = m.mask_cache.density
        ops['nonempty_mask'] = m.nonempty_mask
        ops['tv_value'] = vf.total_variation(m.sdf.grid, m.nonempty_mask)
        # NeuS alpha on random inputs
        n_rays = 16
        vd = T(S.make_rays(n_rays, seed=5)[2])
        rid = T(np.sort(rs.randint(0, n_rays, 300)).astype(np.int64))
        a_sdf = T((rs.standard_normal(300) * 0.05).astype(np.float32))
        a_grad = T(rs.standard_normal((300, 3)).astype(np.float32))
        dist = 0.5 * m.voxel_size
        s_val, alpha = m.neus_alpha_from_sdf_scatter(vd, rid, dist, a_sdf, a_grad, global_step=1500, is_train=True)
        ops.update(alpha_s_val=s_val, alpha=alpha, alpha_s_val_held=m.s_val.data.clone())
        ops['voxel_size'] = m.voxel_size
    # smooth-grad TV regulariser value and its gradient wrt the sdf grid (voxurf_fine.py:412-421)
    m.sdf.grid.grad = None
    m.gradient = m.neus_sdf_gradient()
    tvl = m.density_total_variation(sdf_tv=0, smooth_grad_tv=0.05)
    tvl.backward()
    ops['sgtv_value'] = tvl.detach()
    ops['sgtv_grad'] = m.sdf.grid.grad.clone()
    np.savez_compressed(os.path.join(HERE, 'fine_ops.npz'), **npy(ops))

    # ---------------------------------------------------------------- fine model: full forward + backward
    m = build_fine(vf, sc, ckpt)
    m._set_nonempty_mask()         # nonempty-mask overwrite of the sdf grid, like the real model (voxurf_fine.py:367)
    n_rays = 96
    ro, rd, vd = (T(x) for x in S.make_rays(n_rays, seed=777))
    target = T(S.make_target(vd.numpy()))
    rk = dict(near=0.3, far=6.0, bg=0, stepsize=0.5, render_grad=True, render_depth=True)
    ret = m(ro, rd, vd, global_step=15001, **rk)
    import torch.nn.functional as F
    loss = F.mse_loss(ret['rgb_marched'], target)
    pout = ret['alphainv_cum'][..., -1].clamp(1e-6, 1 - 1e-6)
    loss = loss + 0.001 * (-(pout * torch.log(pout) + (1 - pout) * torch.log(1 - pout)).mean())
    loss = loss + 0.5 * F.mse_loss(ret['rgb_marched0'], target)
    loss.backward()
    out = Snap({k: v for k, v in ret.items() if torch.is_tensor(v)})
    out['s_val'] = ret['s_val']
    out['loss'] = loss
    out['sdf_after_mask'] = m.sdf.grid.data
    out['grad_sdf'] = m.sdf.grid.grad
    out['grad_k0'] = m.k0.grid.grad
    for i, l in enumerate([x for x in m.rgbnet.modules() if isinstance(x, torch.nn.Linear)]):
        out[f'grad_rgbnet_W{i}'] = l.weight.grad
        out[f'grad_rgbnet_b{i}'] = l.bias.grad
    for i, l in enumerate([x for x in m.k_rgbnet.modules() if isinstance(x, torch.nn.Linear)]):
        out[f'grad_k_rgbnet_W{i}'] = l.weight.grad
        out[f'grad_k_rgbnet_b{i}'] = l.bias.grad
    out['full_gradient'] = m.gradient
    # TV add-grad + the trainer's Adam on the sdf grid, one step (run.py:641-659)
    m.sdf_total_variation_add_grad(0.01 * 0.1 / n_rays, True)
    out['grad_sdf_after_tv'] = m.sdf.grid.grad.clone()
    opt = rutils.Adam([{'params': [m.sdf.grid], 'lr': 5e-3}, {'params': [m.k0.grid], 'lr': 1e-1}], betas=(0.9, 0.99))
    opt.step()
    out['sdf_after_adam'] = m.sdf.grid.data
    out['k0_after_adam'] = m.k0.grid.data
    # eval-mode forward (global_step=None keeps the s_val set above)
    with torch.no_grad():
        ret_e = m(ro, rd, vd, **rk)
    out['eval_rgb_marched'] = ret_e['rgb_marched']
    out['eval_normal_marched'] = ret_e['normal_marched']
    out['eval_depth'] = ret_e['depth']
    np.savez_compressed(os.path.join(HERE, 'fine_forward.npz'), **npy(out))

    # ---------------------------------------------------------------- coarse model
    Gc, Cc, Wc = 16, 12, 32
    scc = S.make_coarse_scene(Gc, Cc, Wc, seed=4, mask_G=12)
    ckpt_c = os.path.join(tmp, 'mask_coarse.tar')
    write_mask_ckpt(scc, ckpt_c)
    mc = build_coarse(vc, scc, ckpt_c)
    mc._set_nonempty_mask()
    rkc = dict(near=0.3, far=6.0, bg=0, stepsize=0.5, render_grad=True)
    ret = mc(ro, rd, vd, global_step=2000, **rkc)
    loss = F.mse_loss(ret['rgb_marched'], target)
    loss.backward()
    out = Snap({k: v for k, v in ret.items() if torch.is_tensor(v)})
    out['s_val'] = ret['s_val']
    out['loss'] = loss
    out['sdf_after_mask'] = mc.sdf.grid.data
    out['grad_sdf'] = mc.sdf.grid.grad
    out['grad_k0'] = mc.k0.grid.grad
    for i, l in enumerate([x for x in mc.rgbnet.modules() if isinstance(x, torch.nn.Linear)]):
        out[f'grad_rgbnet_W{i}'] = l.weight.grad
        out[f'grad_rgbnet_b{i}'] = l.bias.grad
    # dense [N,S] formulation pieces (BASELINE config 1)
    with torch.no_grad():
        dp, dmask, dstep = mc.sample_ray_ori(ro[:8], rd[:8], near=0.3, far=1e9, stepsize=0.5, is_train=False)
        out.update(dense_pts=dp, dense_mask=dmask, dense_step=dstep)
        S_ = dp.shape[1]
        dsdf = mc.grid_sampler(dp.reshape(-1, 3), mc.sdf.grid).reshape(8, S_)
        dgrad = mc.grid_sampler(dp.reshape(-1, 3), mc.neus_sdf_gradient(sdf=mc.sdf.grid)).reshape(8, S_, 3)
        _, dalpha = mc.neus_alpha_from_sdf(vd[:8], dstep.repeat(8, 1), dsdf, dgrad, global_step=2000, is_train=True)
        dw, dcum = dvgo.get_ray_marching_ray(dalpha)
        out.update(dense_sdf=dsdf, dense_grad=dgrad, dense_alpha=dalpha, dense_weights=dw, dense_alphainv_cum=dcum)
    np.savez_compressed(os.path.join(HERE, 'coarse_forward.npz'), **npy(out))

    # ---------------------------------------------------------------- utils.Adam, three steps, two lrs
    rs = np.random.RandomState(21)
    p0 = rs.standard_normal(257).astype(np.float32)
    p = torch.nn.Parameter(T(p0).clone())
    opt = rutils.Adam([{'params': [p], 'lr': 5e-3}], betas=(0.9, 0.99))
    traj = []
    for it in range(3):
        g = rs.standard_normal(257).astype(np.float32)
        g[::5] = 0
        p.grad = T(g).clone()
        opt.step()
        traj.append(p.data.clone().numpy())
    st = opt.state[p]
    np.savez_compressed(os.path.join(HERE, 'adam.npz'), traj=np.stack(traj), exp_avg=st['exp_avg'].numpy(),
                        exp_avg_sq=st['exp_avg_sq'].numpy())

    # ---------------------------------------------------------------- MaskGrid (grid.py:212-245)
    mk = rs.uniform(0, 1, (9, 10, 11)) > 0.5
    mg = rgrid.MaskGrid(mask=T(mk), xyz_min=[-1., -1., -1.], xyz_max=[1., 1., 1.])
    q = T(rs.uniform(-1.2, 1.2, (400, 3)).astype(np.float32))
    np.savez_compressed(os.path.join(HERE, 'maskgrid.npz'), out=mg(q).numpy(), scale=mg.xyz2ijk_scale.numpy(),
                        shift=mg.xyz2ijk_shift.numpy())
    print('golden vectors written to', HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == '__main__':
    main()
