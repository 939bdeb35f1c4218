"""Caller side of the path: view rays and training-ray gathering.

Mirrors the module-level helpers of lib/voxurf_fine.py:1001-1175 (the same functions also live in
lib/voxurf_coarse.py:738-910; run.py reaches them as `Model.get_training_rays_*`, run.py:494-529) on three kernels of
the C ABI: vx_rays_of_view (one launch per view instead of ~15 elementwise ATen ops), vx_rays_hit_mask (the
in-mask-cache test without materialising ~600 samples per ray in 64-row chunks) and vx_compact_rows3 (the
`img[mask]` / `rays[mask]` copies behind a device-side row counter: no host sync per view).
Same names, arguments, return values and ordering as the reference.
"""
import time

import numpy as np
import torch

from ._lib import call

_MODES = {'lefttop': 0, 'center': 1, 'random': 2}


def _host_K(K):
    K = K.detach().cpu().numpy() if torch.is_tensor(K) else np.asarray(K)
    return [float(np.float32(v)) for v in K[:3, :3].reshape(-1)]      # tensor - numpy scalar narrows the scalar to fp32


def _host_c2w(c2w):
    c = c2w.detach().cpu().numpy() if torch.is_tensor(c2w) else np.asarray(c2w)
    return [float(np.float32(v)) for v in c[:3, :4].reshape(-1)]


def _device_of(c2w):
    if torch.is_tensor(c2w) and c2w.is_cuda:
        return c2w.device
    return torch.device('cuda', torch.cuda.current_device())


def _rays_of_view(H, W, K, c2w, ndc, inverse_y, flip_x, flip_y, mode='center', ndc_near=1.):
    if mode not in _MODES:
        raise NotImplementedError
    dev = _device_of(c2w)
    H, W = int(H), int(W)
    rays_o = torch.empty(H, W, 3, dtype=torch.float32, device=dev)
    rays_d = torch.empty_like(rays_o)
    viewdirs = torch.empty_like(rays_o)
    ji = jj = None
    if mode == 'random':    # same two draws, in the same order, as i+torch.rand_like(i); j+torch.rand_like(j)
        ji = torch.rand(H, W, dtype=torch.float32, device=dev)
        jj = torch.rand(H, W, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        call('vx_rays_of_view', H, W, _host_K(K), _host_c2w(c2w), int(bool(inverse_y)), int(bool(flip_x)), int(bool(flip_y)),
             _MODES[mode], ji, jj, int(bool(ndc)), float(ndc_near), rays_o, rays_d, viewdirs)
    return rays_o, rays_d, viewdirs


def get_rays(H, W, K, c2w, inverse_y, flip_x, flip_y, mode='center'):
    """lib/voxurf_fine.py:1001-1029"""
    rays_o, rays_d, _ = _rays_of_view(H, W, K, c2w, False, inverse_y, flip_x, flip_y, mode)
    return rays_o, rays_d


def get_rays_np(H, W, K, c2w):
    """lib/voxurf_fine.py:1032-1040 (host numpy helper of the reference; kept as is: it is not on the device path)"""
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing='xy')
    dirs = np.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -np.ones_like(i)], -1)
    rays_d = np.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1)
    rays_o = np.broadcast_to(c2w[:3, 3], np.shape(rays_d))
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """lib/voxurf_fine.py:1043-1062 on explicit rays (elementwise torch; get_rays_of_a_view fuses it into the kernel)."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    o0 = -1. / (W / (2. * focal)) * rays_o[..., 0] / rays_o[..., 2]
    o1 = -1. / (H / (2. * focal)) * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (W / (2. * focal)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = -1. / (H / (2. * focal)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2. * near / rays_o[..., 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


def get_rays_of_a_view(H, W, K, c2w, ndc, inverse_y, flip_x, flip_y, mode='center'):
    """lib/voxurf_fine.py:1065-1070 -> rays_o, rays_d, viewdirs, each (H, W, 3)"""
    return _rays_of_view(H, W, K, c2w, ndc, inverse_y, flip_x, flip_y, mode)


@torch.no_grad()
def get_training_rays(rgb_tr, train_poses, HW, Ks, ndc, inverse_y, flip_x, flip_y):
    """lib/voxurf_fine.py:1073-1095"""
    assert len(np.unique(HW, axis=0)) == 1
    assert len(np.unique(np.asarray(Ks).reshape(len(Ks), -1), axis=0)) == 1
    assert len(rgb_tr) == len(train_poses) and len(rgb_tr) == len(Ks) and len(rgb_tr) == len(HW)
    H, W = HW[0]
    K = Ks[0]
    rays_o_tr = torch.zeros([len(rgb_tr), H, W, 3], device=rgb_tr.device)
    rays_d_tr = torch.zeros([len(rgb_tr), H, W, 3], device=rgb_tr.device)
    viewdirs_tr = torch.zeros([len(rgb_tr), H, W, 3], device=rgb_tr.device)
    imsz = [1] * len(rgb_tr)
    for i, c2w in enumerate(train_poses):
        rays_o, rays_d, viewdirs = get_rays_of_a_view(H=H, W=W, K=K, c2w=c2w, ndc=ndc, inverse_y=inverse_y, flip_x=flip_x,
                                                      flip_y=flip_y)
        rays_o_tr[i].copy_(rays_o)
        rays_d_tr[i].copy_(rays_d)
        viewdirs_tr[i].copy_(viewdirs)
    return rgb_tr, rays_o_tr, rays_d_tr, viewdirs_tr, imsz


@torch.no_grad()
def get_training_rays_flatten(rgb_tr_ori, train_poses, HW, Ks, ndc, inverse_y, flip_x, flip_y):
    """lib/voxurf_fine.py:1098-1124"""
    assert len(rgb_tr_ori) == len(train_poses) and len(rgb_tr_ori) == len(Ks) and len(rgb_tr_ori) == len(HW)
    DEVICE = rgb_tr_ori[0].device
    N = sum(im.shape[0] * im.shape[1] for im in rgb_tr_ori)
    rgb_tr = torch.zeros([N, 3], device=DEVICE)
    rays_o_tr = torch.zeros_like(rgb_tr)
    rays_d_tr = torch.zeros_like(rgb_tr)
    viewdirs_tr = torch.zeros_like(rgb_tr)
    imsz = []
    top = 0
    for c2w, img, (H, W), K in zip(train_poses, rgb_tr_ori, HW, Ks):
        assert img.shape[:2] == (H, W)
        rays_o, rays_d, viewdirs = get_rays_of_a_view(H=H, W=W, K=K, c2w=c2w, ndc=ndc, inverse_y=inverse_y, flip_x=flip_x,
                                                      flip_y=flip_y)
        n = H * W
        rgb_tr[top:top + n].copy_(img.flatten(0, 1))
        rays_o_tr[top:top + n].copy_(rays_o.flatten(0, 1))
        rays_d_tr[top:top + n].copy_(rays_d.flatten(0, 1))
        viewdirs_tr[top:top + n].copy_(viewdirs.flatten(0, 1))
        imsz.append(n)
        top += n
    assert top == N
    return rgb_tr, rays_o_tr, rays_d_tr, viewdirs_tr, imsz


@torch.no_grad()
def get_training_rays_in_maskcache_sampling(rgb_tr_ori, train_poses, HW, Ks, ndc, inverse_y, flip_x, flip_y, model,
                                            render_kwargs, rgbnet_sup_reduce=1):
    """lib/voxurf_fine.py:1127-1164: keep only the rays that hit the coarse geometry (model.hit_coarse_geo), per view in
    pixel order.  One host read at the very end (the per-view counts) instead of one boolean-indexing sync per view."""
    assert len(rgb_tr_ori) == len(train_poses) and len(rgb_tr_ori) == len(Ks) and len(rgb_tr_ori) == len(HW)
    eps_time = time.time()
    DEVICE = rgb_tr_ori[0].device
    N = sum(im.shape[0] * im.shape[1] for im in rgb_tr_ori)
    rgb_tr = torch.zeros([N, 3], device=DEVICE)
    rays_o_tr = torch.zeros_like(rgb_tr)
    rays_d_tr = torch.zeros_like(rgb_tr)
    viewdirs_tr = torch.zeros_like(rgb_tr)
    tops = torch.zeros(len(rgb_tr_ori) + 1, dtype=torch.int64, device=DEVICE)
    for k, (c2w, img, (H, W), K) in enumerate(zip(train_poses, rgb_tr_ori, HW, Ks)):
        assert img.shape[:2] == (H, W)
        rays_o, rays_d, viewdirs = get_rays_of_a_view(H=H, W=W, K=K, c2w=c2w, ndc=ndc, inverse_y=inverse_y, flip_x=flip_x,
                                                      flip_y=flip_y)
        mask = model.hit_coarse_geo(rays_o=rays_o, rays_d=rays_d, **render_kwargs).reshape(-1)
        incl = torch.cumsum(mask, 0, dtype=torch.int32)
        call('vx_compact_rows3', mask, incl, mask.numel(), tops[k:k + 1], tops[k + 1:k + 2], N,
             img.reshape(-1, 3).float().contiguous(), rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), viewdirs.reshape(-1, 3),
             rgb_tr, rays_o_tr, rays_d_tr, viewdirs_tr)
    tops_host = tops.cpu()
    top = int(tops_host[-1])
    imsz = [t for t in (tops_host[1:] - tops_host[:-1])]
    print('get_training_rays_in_maskcache_sampling: ratio', top / max(N, 1))
    print('get_training_rays_in_maskcache_sampling: finish (eps time:', time.time() - eps_time, 'sec)')
    return rgb_tr[:top], rays_o_tr[:top], rays_d_tr[:top], viewdirs_tr[:top], imsz


def batch_indices_generator(N, BS):
    """lib/voxurf_fine.py:1167-1175"""
    idx, top = torch.LongTensor(np.random.permutation(N)), 0
    while True:
        if top + BS > N:
            idx, top = torch.LongTensor(np.random.permutation(N)), 0
        yield idx[top:top + BS]
        top += BS
