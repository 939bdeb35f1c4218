"""TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by voxurf_b200/.

Runs the REFERENCE'S OWN Python model code (lib/voxurf_fine.py, lib/voxurf_coarse.py, lib/grid.py, lib/utils.py) unmodified
on a GPU, from `baseline/_ref/lib` -- a git-ignored copy of /root/reference/lib that __graft_entry__.build() places there in
the build container (it travels to the GPU box with the snapshot, like oracle/_ref; /root/reference itself does not exist
there).  Two back ends for the extension modules the reference JIT-loads with torch.utils.cpp_extension.load:

  backend='ref'   the reference's own CUDA kernels, compiled unmodified into oracle/_ref by oracle/build_ref.py, plus
                  torch_scatter.segment_coo as out.index_add_ (SURVEY.md 8c "GPU oracle"): the reference's GPU path
                   -- what this repository has to beat, timed on the same B200;
  backend='b200'  voxurf_b200's shim modules of the same names (render_utils_cuda, total_variation_cuda, torch_scatter):
                  the reference model running unmodified on this repository's C ABI ("integration
                  oracle").

Missing third-party imports of the reference (cv2, matplotlib, mcubes, plyfile, imageio, skimage, trimesh) are stubbed;
none is on the render path.  `train_step` restates the loop body of run.py:600-683 for the surf stages around the
reference's own model / optimizer objects.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_COPY = os.path.join(ROOT, 'baseline', '_ref')


def available():
    return os.path.isfile(os.path.join(REF_COPY, 'lib', 'voxurf_fine.py'))


def place_copy(src='/root/reference'):
    """build container only: copy the reference's lib/*.py (not its CUDA sources: those are compiled where they lie by
    oracle/build_ref.py) to baseline/_ref/lib.  Git-ignored; not product source."""
    import shutil
    if not os.path.isdir(os.path.join(src, 'lib')):
        return False
    dst = os.path.join(REF_COPY, 'lib')
    os.makedirs(dst, exist_ok=True)
    for f in os.listdir(os.path.join(src, 'lib')):
        if f.endswith('.py'):
            shutil.copyfile(os.path.join(src, 'lib', f), os.path.join(dst, f))
    return True


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Anything(self.__name__ + '.' + name)

    def __call__(self, *a, **k):
        return None


def load_lib(backend, extra=()):
    """-> namespace with the reference modules voxurf_fine, voxurf_coarse, grid, utils (+ `extra`, e.g.
    voxurf_womask_fine) imported against `backend`."""
    assert backend in ('ref', 'b200')
    assert available(), 'baseline/_ref/lib is missing: run __graft_entry__.build() in the build container'
    for name in ['cv2', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.cm', 'mcubes', 'plyfile', 'imageio', 'skimage', 'skimage.measure',
                 'trimesh']:
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _Anything(name)
    ts = types.ModuleType('torch_scatter')
    if backend == 'ref':
        def segment_coo(src, index, out=None, reduce='sum'):
            assert reduce == 'sum'
            return out.index_add_(0, index, src)
        ts.segment_coo = segment_coo
        from oracle import build_ref
        mods = {n: build_ref.load_ref(n) for n in ('render_utils_cuda', 'total_variation_cuda')}
        assert all(m is not None for m in mods.values()), 'oracle/_ref/*.so missing: run oracle/build_ref.py in the build container'
        mods['ub360_utils_cuda'] = build_ref.load_ref('ub360_utils_cuda')     # only the womask models load it
    else:
        from voxurf_b200 import render_utils_cuda, total_variation_cuda, torch_scatter, ub360_utils_cuda
        ts.segment_coo = torch_scatter.segment_coo
        mods = {'render_utils_cuda': render_utils_cuda, 'total_variation_cuda': total_variation_cuda, 'ub360_utils_cuda': ub360_utils_cuda}
    saved_ts = sys.modules.get('torch_scatter')
    sys.modules['torch_scatter'] = ts
    import torch.utils.cpp_extension as ce
    saved_load = ce.load
    ce.load = lambda name, **kw: mods[name]
    for k in [k for k in sys.modules if k == 'lib' or k.startswith('lib.')]:
        del sys.modules[k]
    sys.path.insert(0, REF_COPY)
    try:
        ns = types.SimpleNamespace(backend=backend)
        for m in ('grid', 'utils', 'dvgo_ori', 'voxurf_fine', 'voxurf_coarse') + tuple(extra):
            setattr(ns, m, importlib.import_module('lib.' + m))
    finally:
        sys.path.remove(REF_COPY)
        ce.load = saved_load
        if saved_ts is not None:
            sys.modules['torch_scatter'] = saved_ts
        else:
            del sys.modules['torch_scatter']
        for k in [k for k in sys.modules if k == 'lib' or k.startswith('lib.')]:
            del sys.modules[k]      # the next load_lib() imports a fresh copy against its own back end
    return ns


class Cfg(dict):
    """attribute-style config like the mmcv Config objects run.py passes around"""
    __getattr__ = dict.__getitem__


def write_mask_ckpt(path, density, act_shift, voxel_size_ratio=1.0):
    torch.save({'MaskCache_kwargs': {'xyz_min': [-1., -1., -1.], 'xyz_max': [1., 1., 1.], 'act_shift': act_shift,
                                     'voxel_size_ratio': voxel_size_ratio, 'nearest': False},
                'model_state_dict': {'density': torch.as_tensor(density).cpu()}}, path)


class cuda_default:
    """run.py:954 makes CUDA float tensors the default; the reference's constructors rely on it"""

    def __enter__(self):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            torch.set_default_tensor_type('torch.cuda.FloatTensor')
        return self

    def __exit__(self, *a):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            torch.set_default_tensor_type('torch.FloatTensor')


def build_fine(ns, G, C, width, mask_ckpt, cfg_model, sdf, k0, mlps, smooth_ksize=0):
    """Reference fine model (lib/voxurf_fine.py:25-203) with given parameters (CPU or CUDA tensors)."""
    vf = ns.voxurf_fine
    with cuda_default():
        m = vf.Voxurf(xyz_min=torch.tensor([-1., -1., -1.]).cuda(), xyz_max=torch.tensor([1., 1., 1.]).cuda(), num_voxels=G ** 3, num_voxels_base=G ** 3,
                      mask_cache_path=mask_ckpt, rgbnet_dim=C, rgbnet_width=width, smooth_ksize=smooth_ksize, smooth_sigma=0.8,
                      **cfg_model).cuda()
        assert tuple(int(w) for w in m.world_size) == (G, G, G), m.world_size
        with torch.no_grad():
            m.sdf.grid.data.copy_(torch.as_tensor(sdf).cuda())
            m.k0.grid.data.copy_(torch.as_tensor(k0).cuda())
            for seq, layers in ((m.rgbnet, mlps[0]), (m.k_rgbnet, mlps[1])):
                lin = [x for x in seq.modules() if isinstance(x, torch.nn.Linear)]
                for l, (W, b) in zip(lin, layers):
                    l.weight.data.copy_(torch.as_tensor(W).cuda()); l.bias.data.copy_(torch.as_tensor(b).cuda())
            if m.mask_cache is not None:
                m._set_nonempty_mask()
    return m


def make_optimizer(ns, model, cfg_train):
    with cuda_default():
        return ns.utils.create_optimizer_or_freeze_model(model, cfg_train, global_step=0)


def train_step(model, optimizer, cfg_train, render_kwargs, batch, global_step):
    """run.py:600-683 for a surf stage (fine: ori_tv False; coarse: ori_tv True).  -> (loss, render_result)"""
    rays_o, rays_d, viewdirs, target = batch
    c = cfg_train
    with cuda_default():
        rr = model(rays_o, rays_d, viewdirs, global_step=global_step, **render_kwargs)
        optimizer.zero_grad(set_to_none=True)
        loss = c.weight_main * F.mse_loss(rr['rgb_marched'], target)
        if c.weight_entropy_last > 0:
            pout = rr['alphainv_cum'][..., -1].clamp(1e-6, 1 - 1e-6)
            loss = loss + c.weight_entropy_last * (-(pout * torch.log(pout) + (1 - pout) * torch.log(1 - pout)).mean())
        tv_iter = c.tv_from < global_step < c.tv_end and global_step % c.tv_every == 0
        if tv_iter and c.weight_tv_density > 0:
            tv = c.tv_terms
            if tv['smooth_grad_tv'] > 0:
                loss = loss + c.weight_tv_density * model.density_total_variation(sdf_tv=0, smooth_grad_tv=tv['smooth_grad_tv'])
            if c.get('ori_tv', False):
                loss = loss + c.weight_tv_density * model.density_total_variation(sdf_tv=tv['sdf_tv'], smooth_grad_tv=0)
                if c.weight_tv_k0 > 0:
                    loss = loss + c.weight_tv_k0 * model.k0_total_variation()
        if c.get('weight_rgb0', 0.) > 0:
            loss = loss + F.mse_loss(rr['rgb_marched0'], target) * c.weight_rgb0
        loss.backward()
        if tv_iter and not c.get('ori_tv', False):
            if c.weight_tv_density > 0 and c.tv_terms['sdf_tv'] > 0:
                model.sdf_total_variation_add_grad(c.weight_tv_density * c.tv_terms['sdf_tv'] / len(rays_o), global_step < c.tv_dense_before)
            if c.weight_tv_k0 > 0:
                model.k0_total_variation_add_grad(c.weight_tv_k0 / len(rays_o), global_step < c.tv_dense_before)
        optimizer.step()
        decay = 0.1 ** (1 / (c.lrate_decay * 1000))
        for g in optimizer.param_groups:
            g['lr'] = g['lr'] * decay
    return loss.detach(), rr
