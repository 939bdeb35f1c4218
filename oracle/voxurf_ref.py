"""TEST INFRASTRUCTURE ONLY -- CPU oracle.  Never imported by voxurf_b200/.

fp32 CPU-torch restatement of the Python half of Voxurf's ray-batch render path: everything
`lib/voxurf_fine.py`, `lib/voxurf_coarse.py`, `lib/grid.py`, `lib/utils.py` and `run.py` do between
the native operators.  Functional style (plain tensors in, tensors out); every function cites the
reference lines (relative to /root/reference) it follows.  The native operators themselves are
restated in C (oracle/ref_kernels.c, wrapped by oracle/kernels.py).

Pinning: tests/golden/make_golden.py imports the *reference's own* Python modules (with the
missing third-party imports stubbed) and stores their CPU outputs for the functions below;
tests/test_oracle_golden.py checks this file against those vectors.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import kernels as K


# ------------------------------------------------------------------------------------------------
# grids
# ------------------------------------------------------------------------------------------------
def ind_norm(xyz, xyz_min, xyz_max):
    """grid.py:52 / voxurf_fine.py:516 -- world xyz -> grid_sample coords (z,y,x order, [-1,1])."""
    return ((xyz.reshape(1, 1, 1, -1, 3) - xyz_min) / (xyz_max - xyz_min)).flip((-1,)) * 2 - 1


def grid_trilinear(grid, xyz, xyz_min, xyz_max, mode='bilinear'):
    """grid.py:47-58 (DenseGrid.forward), voxurf_fine.py:516-519, voxurf_coarse.py:435-452.
    grid (1,C,X,Y,Z), xyz (P,3) -> (P,C), zeros padding, align_corners=True."""
    out = F.grid_sample(grid, ind_norm(xyz, xyz_min, xyz_max), mode=mode, align_corners=True)
    return out.reshape(grid.shape[1], -1).T


def dense_grid_forward(grid, xyz, xyz_min, xyz_max):
    """grid.py:47-58: (P,C), squeezed to (P,) when C == 1."""
    out = grid_trilinear(grid, xyz, xyz_min, xyz_max)
    return out.squeeze(-1) if grid.shape[1] == 1 else out


def mask_cache_density(raw_density, ks=3):
    """voxurf_fine.py:924-925: 3^3 max-pool of the coarse density at construction."""
    return F.max_pool3d(raw_density, kernel_size=ks, padding=ks // 2, stride=1)


def mask_cache_forward(density, xyz, mc_xyz_min, mc_xyz_max, act_shift, voxel_size_ratio, thres, nearest=False):
    """voxurf_fine.py:930-942 (== voxurf_coarse.py:684-696). -> bool (P,)"""
    d = F.grid_sample(density, ind_norm(xyz, mc_xyz_min, mc_xyz_max), align_corners=True,
                      mode='nearest' if nearest else 'bilinear')
    alpha = 1 - torch.exp(-F.softplus(d + act_shift) * voxel_size_ratio)
    return alpha.reshape(-1) >= thres


def mask_grid_params(mask_shape, xyz_min, xyz_max):
    """grid.py:230-232 (MaskGrid): xyz2ijk_scale / shift."""
    scale = (torch.tensor(list(mask_shape), dtype=torch.float32) - 1) / (xyz_max - xyz_min)
    return scale, -xyz_min * scale


def gaussian_kernel3d(ksize, sigma):
    """voxurf_fine.py:246-254: normalised exp(-r^2/2sigma^2), float32 (ksize,ksize,ksize)."""
    r = np.arange(-(ksize // 2), ksize // 2 + 1, 1)
    xx, yy, zz = np.meshgrid(r, r, r)
    k = np.exp(-(xx ** 2 + yy ** 2 + zz ** 2) / (2 * sigma ** 2))
    k = torch.from_numpy(k).float()
    return k / k.sum()


def conv3d_replicate(x, kernel):
    """voxurf_fine.py:253: Conv3d(1,1,k,padding=k//2,padding_mode='replicate'), zero bias.
    x (1,1,X,Y,Z) or (B,1,X,Y,Z)."""
    p = kernel.shape[-1] // 2
    return F.conv3d(F.pad(x, (p,) * 6, mode='replicate'), kernel[None, None])


def tv_smooth_kernel():
    """voxurf_fine.py:205-239 with sigma=0: the binomial 3^3 kernel / its sum (tv_smooth_conv)."""
    k = np.asarray([[[1, 2, 1], [2, 4, 2], [1, 2, 1]],
                    [[2, 4, 2], [4, 8, 4], [2, 4, 2]],
                    [[1, 2, 1], [2, 4, 2], [1, 2, 1]]], dtype=np.float64)
    return torch.from_numpy(k / k.sum()).float()


def sdf_gradient_grid(sdf, voxel_size, mode='interpolate'):
    """voxurf_fine.py:440-460 (== voxurf_coarse.py:323-346). sdf (1,1,X,Y,Z) -> (1,3,X,Y,Z).
    'interpolate': central difference in the interior, zero on the two boundary planes of each axis."""
    g = torch.zeros([1, 3, *sdf.shape[-3:]], dtype=sdf.dtype)
    if mode == 'interpolate':
        g[:, 0, 1:-1, :, :] = (sdf[:, 0, 2:, :, :] - sdf[:, 0, :-2, :, :]) / 2 / voxel_size
        g[:, 1, :, 1:-1, :] = (sdf[:, 0, :, 2:, :] - sdf[:, 0, :, :-2, :]) / 2 / voxel_size
        g[:, 2, :, :, 1:-1] = (sdf[:, 0, :, :, 2:] - sdf[:, 0, :, :, :-2]) / 2 / voxel_size
    elif mode == 'raw':
        g[:, 0, :-1, :, :] = (sdf[:, 0, 1:, :, :] - sdf[:, 0, :-1, :, :]) / voxel_size
        g[:, 1, :, :-1, :] = (sdf[:, 0, :, 1:, :] - sdf[:, 0, :, :-1, :]) / voxel_size
        g[:, 2, :, :, :-1] = (sdf[:, 0, :, :, 1:] - sdf[:, 0, :, :, :-1]) / voxel_size
    else:
        raise NotImplementedError(mode)
    return g


def sample_sdfs(xyz, grid, displace_list, xyz_min, xyz_max, voxel_size, use_grad_norm=False):
    """voxurf_fine.py:537-577.  For every displacement d and axis a (in z,y,x order) sample the grid at
    index-space offsets -d/+d (indices clamped to the grid, :554-556), finite-difference by the
    *clamped* index distance and by voxel_size (:562-566).
    -> feat (P, 6*L) [index (a*2+sign)*L + l], grad (P, 3*L) [index a*L + l], axis order z,y,x."""
    P = xyz.shape[0]
    L = len(displace_list)
    n = ind_norm(xyz, xyz_min, xyz_max)
    gs = grid.shape[-3:]
    size_zyx = torch.tensor([gs[2], gs[1], gs[0]])
    ind = ((n + 1) / 2) * (size_zyx - 1)
    offset = torch.tensor([[-1, 0, 0], [1, 0, 0], [0, -1, 0], [0, 1, 0], [0, 0, -1], [0, 0, 1]])
    displace = torch.tensor(displace_list)
    offset = offset[:, None, :] * displace[None, :, None]
    all_ind = (ind.unsqueeze(-2) + offset.view(-1, 3)).view(1, 1, 1, -1, 3)
    all_ind[..., 0] = all_ind[..., 0].clamp(min=0, max=size_zyx[0] - 1)
    all_ind[..., 1] = all_ind[..., 1].clamp(min=0, max=size_zyx[1] - 1)
    all_ind[..., 2] = all_ind[..., 2].clamp(min=0, max=size_zyx[2] - 1)
    all_ind_norm = (all_ind / (size_zyx - 1)) * 2 - 1
    feat = F.grid_sample(grid, all_ind_norm, mode='bilinear', align_corners=True)
    all_ind = all_ind.view(1, 1, 1, -1, 6, L, 3)
    diff = all_ind[:, :, :, :, 1::2, :, :] - all_ind[:, :, :, :, 0::2, :, :]
    diff, _ = diff.max(dim=-1)
    feat_ = feat.view(1, 1, 1, -1, 6, L)
    feat_diff = feat_[:, :, :, :, 1::2, :] - feat_[:, :, :, :, 0::2, :]
    grad = feat_diff / diff / voxel_size
    feat = feat.view(P, 6, L)
    grad = grad.view(P, 3, L)
    if use_grad_norm:
        grad = grad / (grad.norm(dim=1, keepdim=True) + 1e-5)
    return feat.reshape(P, 6 * L), grad.reshape(P, 3 * L)


def fine_grid_sampler(xyz, sdf_grid, xyz_min, xyz_max, voxel_size):
    """voxurf_fine.py:502-534 with sample_ret=True, sample_grad=True: -> sdf (P,), grad (P,3) in
    x,y,z order, feat (P,6) reordered x-,x+,y-,y+,z-,z+ (:525-526)."""
    sdf = grid_trilinear(sdf_grid, xyz, xyz_min, xyz_max).squeeze(-1)
    feat, grad = sample_sdfs(xyz, sdf_grid, [1.0], xyz_min, xyz_max, voxel_size, use_grad_norm=False)
    feat = torch.cat([feat[:, 4:6], feat[:, 2:4], feat[:, 0:2]], dim=-1)
    grad = torch.cat([grad[:, [2]], grad[:, [1]], grad[:, [0]]], dim=-1)
    return sdf, grad, feat


# ------------------------------------------------------------------------------------------------
# NeuS alpha, transmittance, compositing
# ------------------------------------------------------------------------------------------------
def s_val_schedule(global_step, s_ratio, s_start, step_start=0):
    """voxurf_fine.py:468."""
    return 1. / (global_step + s_ratio / s_start - step_start) * s_ratio


def neus_alpha_from_sdf_scatter(viewdirs, ray_id, dist, sdf, gradients, s_val):
    """voxurf_fine.py:463-500 (== voxurf_coarse.py:348-382), use_mid=True, cos_anneal_ratio=1.
    s_val is the float32 value held in self.s_val (:469)."""
    dirs = viewdirs[ray_id]
    s = torch.ones(1) * s_val
    inv_s = torch.ones(1) / s
    true_cos = (dirs * gradients).sum(-1, keepdim=True)
    cos_anneal_ratio = 1.0
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio) + F.relu(-true_cos) * cos_anneal_ratio)
    sdf = sdf.unsqueeze(-1)
    dist = torch.as_tensor(dist, dtype=torch.float32)
    est_next = sdf + iter_cos * dist.reshape(-1, 1) * 0.5
    est_prev = sdf - iter_cos * dist.reshape(-1, 1) * 0.5
    prev_cdf = torch.sigmoid(est_prev * inv_s.reshape(-1, 1))
    next_cdf = torch.sigmoid(est_next * inv_s.reshape(-1, 1))
    p = prev_cdf - next_cdf
    c = prev_cdf
    return ((p + 1e-5) / (c + 1e-5)).clip(0.0, 1.0).squeeze(-1)


class Alphas2Weights(torch.autograd.Function):
    """voxurf_fine.py:981-997 on top of the C restatement of render_utils_kernel.cu:576-707."""

    @staticmethod
    def forward(ctx, alpha, ray_id, N):
        weights, T, alphainv_last, i_start, i_end = K.alpha2weight(alpha, ray_id, N)
        if alpha.requires_grad:
            ctx.save_for_backward(alpha, weights, T, alphainv_last, i_start, i_end)
            ctx.n_rays = N
        return weights, alphainv_last

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_weights, grad_last):
        alpha, weights, T, alphainv_last, i_start, i_end = ctx.saved_tensors
        grad = K.alpha2weight_backward(alpha, weights, T, alphainv_last, i_start, i_end, ctx.n_rays,
                                       grad_weights, grad_last)
        return grad, None, None


def segment_coo(src, index, out):
    """torch_scatter.segment_coo(src, index, out=zeros, reduce='sum') with a sorted index
    (voxurf_fine.py:753-777): out[index[i]] += src[i], accumulated in index order."""
    return out.index_add_(0, index, src)


def positional_encoding(x, freqs):
    """voxurf_fine.py:694-698: [x, sin(x*f), cos(x*f)] with the (dim, freq) flattening of the reference."""
    emb = (x.unsqueeze(-1) * freqs).flatten(-2)
    return torch.cat([x, emb.sin(), emb.cos()], -1)


def mlp(x, layers):
    """nn.Sequential(Linear, ReLU, ..., Linear) of voxurf_fine.py:132-139. layers = [(W,b), ...]."""
    for i, (W, b) in enumerate(layers):
        x = F.linear(x, W, b)
        if i + 1 < len(layers):
            x = F.relu(x)
    return x


# ------------------------------------------------------------------------------------------------
# the two model forwards
# ------------------------------------------------------------------------------------------------
def sample_ray(rays_o, rays_d, xyz_min, xyz_max, near, stepsize, voxel_size):
    """voxurf_fine.py:593-617 (== voxurf_coarse.py:454-486 with maskout=True, use_bg=False)."""
    far = 1e9
    stepdist = float(stepsize * voxel_size)
    pts, mask_outbbox, ray_id, step_id, n_steps, t_min, t_max = K.sample_pts_on_rays(
        rays_o.contiguous(), rays_d.contiguous(), xyz_min, xyz_max, near, far, stepdist)
    n_steps = ray_id.unique(return_counts=True)[1]
    inb = ~mask_outbbox
    return pts[inb], ray_id[inb], step_id[inb], mask_outbbox, n_steps


def fine_forward(m, rays_o, rays_d, viewdirs, global_step=None, *, near, stepsize, bg, render_grad=False,
                 render_depth=False):
    """voxurf_fine.py:620-802.  `m` is a dict/namespace with: xyz_min, xyz_max, voxel_size (0-d tensors),
    sdf (1,1,X,Y,Z), k0 (1,C,X,Y,Z), rgbnet / k_rgbnet = [(W,b),...], posfreq, viewfreq, k_posfreq,
    k_viewfreq, grad_feat (== sdf_feat, tuple), use_grad_norm, center_sdf, k_center_sdf, k_res,
    fast_color_thres, s_ratio, s_start, step_start, s_val (float, used when global_step is None),
    smooth_kernel (or None), mask_cache (dict or None)."""
    N = len(rays_o)
    ray_pts, ray_id, step_id, mask_outbbox, n_steps = sample_ray(
        rays_o, rays_d, m['xyz_min'], m['xyz_max'], near, stepsize, m['voxel_size'])
    mc = m.get('mask_cache')
    if mc is not None:  # :631-636
        mask = mask_cache_forward(mc['density'], ray_pts, mc['xyz_min'], mc['xyz_max'], mc['act_shift'],
                                  mc['voxel_size_ratio'], mc['thres'])
        ray_pts, ray_id, step_id = ray_pts[mask], ray_id[mask], step_id[mask]
        mask_outbbox[~mask_outbbox] |= ~mask
    sdf_grid = conv3d_replicate(m['sdf'], m['smooth_kernel']) if m.get('smooth_kernel') is not None else m['sdf']
    sdf, gradient, feat = fine_grid_sampler(ray_pts, sdf_grid, m['xyz_min'], m['xyz_max'], m['voxel_size'])
    dist = stepsize * m['voxel_size']
    if global_step is not None:  # :466-473
        s_val = s_val_schedule(global_step, m['s_ratio'], m['s_start'], m.get('step_start', 0))
        s_val_held = float(torch.ones(1) * s_val)
        ret_s_val = s_val
    else:
        s_val_held = m['s_val']
        ret_s_val = 0
    alpha = neus_alpha_from_sdf_scatter(viewdirs, ray_id, dist, sdf, gradient, s_val_held)
    mask = None
    thres = m['fast_color_thres']
    if thres > 0:  # :647-654
        mask = alpha > thres
        alpha, ray_id, ray_pts, step_id, gradient, sdf = (t[mask] for t in (alpha, ray_id, ray_pts, step_id, gradient, sdf))
    weights, alphainv_last = Alphas2Weights.apply(alpha, ray_id, N)
    if thres > 0:  # :668-676 (weights are NOT recomputed)
        mask = weights > thres
        weights, alpha, ray_pts, ray_id, step_id, gradient, sdf = (
            t[mask] for t in (weights, alpha, ray_pts, ray_id, step_id, gradient, sdf))
    k0 = dense_grid_forward(m['k0'], ray_pts, m['xyz_min'], m['xyz_max'])
    disp = sorted(set(m['grad_feat']))
    all_feat, all_grad = (sample_sdfs(ray_pts, sdf_grid, disp, m['xyz_min'], m['xyz_max'], m['voxel_size'],
                                      use_grad_norm=m['use_grad_norm']) if len(disp) > 0 else (None, None))
    full_gradient = sdf_gradient_grid(m['sdf'], m['voxel_size'])  # :692
    viewdirs_emb = positional_encoding(viewdirs, m['viewfreq'])
    rays_xyz = (ray_pts - m['xyz_min']) / (m['xyz_max'] - m['xyz_min'])
    xyz_emb = positional_encoding(rays_xyz, m['posfreq'])
    rgb_feat = torch.cat([xyz_emb, viewdirs_emb[ray_id]], -1)  # use_rgbnet_k0=False (:704-707)
    hier = []
    if m['center_sdf']:
        hier.append(sdf[:, None])
    if len(disp) > 0:
        hier += [all_feat, all_grad]
    if hier:
        rgb_feat = torch.cat([rgb_feat, *hier], dim=-1)
    rgb_logit = mlp(rgb_feat, m['rgbnet'])
    rgb = torch.sigmoid(rgb_logit)
    # second network (:722-756), k_detach_1 = k_detach_2 = True
    k_xyz_emb = positional_encoding(rays_xyz, m['k_posfreq'])
    k_viewdirs_emb = positional_encoding(viewdirs, m['k_viewfreq'])
    k_feat = torch.cat([k0, k_xyz_emb, k_viewdirs_emb[ray_id]], -1)
    extra = [gradient]
    if m.get('k_center_sdf', False):
        extra.append(sdf[:, None])
    k_feat = torch.cat([k_feat, *extra], dim=-1)
    if m.get('k_res', True):
        k_feat = torch.cat([k_feat, rgb_logit.detach()], dim=-1)
    k_rgb_logit = rgb_logit.detach() + mlp(k_feat, m['k_rgbnet'])
    k_rgb = torch.sigmoid(k_rgb_logit)
    k_rgb_marched = segment_coo(weights.unsqueeze(-1) * k_rgb, ray_id, torch.zeros([N, 3])) + alphainv_last[..., None] * bg
    k_rgb_marched = k_rgb_marched.clamp(0, 1)
    rgb_marched = segment_coo(weights.unsqueeze(-1) * rgb, ray_id, torch.zeros([N, 3])) + alphainv_last[..., None] * bg
    normal_marched = None
    if render_grad:  # :765-769
        normal = gradient / (gradient.norm(2, -1, keepdim=True) + 1e-6)
        normal_marched = segment_coo(weights.unsqueeze(-1) * normal, ray_id, torch.zeros([N, 3]))
    depth, disp_out = None, 0
    if render_depth:  # :773-777
        with torch.no_grad():
            depth = segment_coo(weights * step_id * dist, ray_id, torch.zeros([N]))
        disp_out = 1 / depth
    return {
        'alphainv_cum': alphainv_last, 'weights': weights, 'rgb_marched': k_rgb_marched,
        'rgb_marched0': rgb_marched, 'normal_marched': normal_marched, 'raw_alpha': alpha, 'raw_rgb': rgb,
        'depth': depth, 'disp': disp_out, 'mask': mask, 'mask_outbbox': mask_outbbox, 'gradient': gradient,
        'gradient_error': None, 's_val': ret_s_val,
        # extras for stage-wise parity checks (not in the reference's ret_dict)
        '_ray_id': ray_id, '_step_id': step_id, '_ray_pts': ray_pts, '_sdf': sdf, '_k0': k0,
        '_full_gradient': full_gradient, '_n_steps': n_steps,
    }


def coarse_forward(m, rays_o, rays_d, viewdirs, global_step=None, *, near, stepsize, bg, render_grad=False):
    """voxurf_coarse.py:513-619.  Differences from the fine model: SDF sampled from the smoothed grid but the
    gradient trilinearly sampled from the grid-level central difference of the RAW grid (:531-535); one
    weights>thres compaction followed by a second alpha2weight (:540-550); single rgbnet fed
    [k0, xyz_emb, view_emb, normal] (:552-571); background blended by 1-sum(w) and clamped (:575-583)."""
    N = len(rays_o)
    ray_pts, ray_id, step_id, mask_outbbox, n_steps = sample_ray(
        rays_o, rays_d, m['xyz_min'], m['xyz_max'], near, stepsize, m['voxel_size'])
    mc = m.get('mask_cache')
    mask = None
    if mc is not None:
        mask = mask_cache_forward(mc['density'], ray_pts, mc['xyz_min'], mc['xyz_max'], mc['act_shift'],
                                  mc['voxel_size_ratio'], mc['thres'])
        ray_pts, ray_id, step_id = ray_pts[mask], ray_id[mask], step_id[mask]
        mask_outbbox[~mask_outbbox] |= ~mask
    sdf_grid = conv3d_replicate(m['sdf'], m['smooth_kernel']) if m.get('smooth_kernel') is not None else m['sdf']
    sdf = grid_trilinear(sdf_grid, ray_pts, m['xyz_min'], m['xyz_max']).squeeze(-1)
    full_gradient = sdf_gradient_grid(m['sdf'], m['voxel_size'])
    gradient = grid_trilinear(full_gradient, ray_pts, m['xyz_min'], m['xyz_max'])
    dist = stepsize * m['voxel_size']
    if global_step is not None:
        s_val = s_val_schedule(global_step, m['s_ratio'], m['s_start'], m.get('step_start', 0))
        s_val_held = float(torch.ones(1) * s_val)
        ret_s_val = s_val
    else:
        s_val_held = m['s_val']
        ret_s_val = 0
    alpha = neus_alpha_from_sdf_scatter(viewdirs, ray_id, dist, sdf, gradient, s_val_held)
    weights, alphainv_last = Alphas2Weights.apply(alpha, ray_id, N)
    thres = m['fast_color_thres']
    if thres > 0:
        mask = weights > thres
        ray_pts, ray_id, step_id, alpha, gradient = (t[mask] for t in (ray_pts, ray_id, step_id, alpha, gradient))
    weights, alphainv_last = Alphas2Weights.apply(alpha, ray_id, N)
    k0 = dense_grid_forward(m['k0'], ray_pts, m['xyz_min'], m['xyz_max'])
    rays_xyz = (ray_pts - m['xyz_min']) / (m['xyz_max'] - m['xyz_min'])
    xyz_emb = positional_encoding(rays_xyz, m['posfreq'])
    viewdirs_emb = positional_encoding(viewdirs, m['viewfreq'])
    rgb_feat = torch.cat([k0, xyz_emb, viewdirs_emb[ray_id]], -1)
    normal = gradient / (gradient.norm(dim=-1, keepdim=True) + 1e-5)
    rgb_feat = torch.cat([rgb_feat, normal], -1)
    rgb_logit = mlp(rgb_feat, m['rgbnet'])
    rgb = torch.sigmoid(rgb_logit)
    rgb_marched = segment_coo(weights.unsqueeze(-1) * rgb, ray_id, torch.zeros([N, 3]))
    cum_weights = segment_coo(weights.unsqueeze(-1), ray_id, torch.zeros([N, 1]))
    rgb_marched = (rgb_marched + (1 - cum_weights) * bg).clamp(0, 1)
    normal_marched = None
    if render_grad:
        nrm = gradient / (gradient.norm(2, -1, keepdim=True) + 1e-6)
        normal_marched = segment_coo(weights.unsqueeze(-1) * nrm, ray_id, torch.zeros([N, 3]))
    return {
        'alphainv_cum': alphainv_last, 'weights': weights, 'rgb_marched': rgb_marched,
        'normal_marched': normal_marched, 'raw_alpha': alpha, 'raw_rgb': rgb, 'depth': None, 'disp': None,
        'mask': mask, 'mask_outbbox': mask_outbbox, 'gradient': gradient, 'gradient_error': None,
        's_val': ret_s_val,
        '_ray_id': ray_id, '_step_id': step_id, '_ray_pts': ray_pts, '_sdf': sdf, '_k0': k0,
        '_full_gradient': full_gradient, '_n_steps': n_steps,
    }


# ------------------------------------------------------------------------------------------------
# the reference's dense [N,S] PyTorch formulation (BASELINE.json config 1; CPU baseline)
# ------------------------------------------------------------------------------------------------
def sample_ray_ori(rays_o, rays_d, xyz_min, xyz_max, grid_shape, near, far, stepsize, voxel_size):
    """voxurf_coarse.py:488-510 (is_train=False: no jitter).  -> pts (N,S,3), mask_outbbox (N,S), step (1,S)"""
    n_samples = int(np.linalg.norm(np.array(grid_shape) + 1) / stepsize) + 1
    vec = torch.where(rays_d == 0, torch.full_like(rays_d, 1e-6), rays_d)
    rate_a = (xyz_max - rays_o) / vec
    rate_b = (xyz_min - rays_o) / vec
    t_min = torch.minimum(rate_a, rate_b).amax(-1).clamp(min=near, max=far)
    t_max = torch.maximum(rate_a, rate_b).amin(-1).clamp(min=near, max=far)
    mask_outbbox = (t_max <= t_min)
    rng = torch.arange(n_samples)[None].float()
    step = stepsize * voxel_size * rng
    interpx = (t_min[..., None] + step / rays_d.norm(dim=-1, keepdim=True))
    pts = rays_o[..., None, :] + rays_d[..., None, :] * interpx[..., None]
    mask_outbbox = mask_outbbox[..., None] | ((xyz_min > pts) | (pts > xyz_max)).any(dim=-1)
    return pts, mask_outbbox, step


def neus_alpha_from_sdf_dense(viewdirs, steps, sdf, gradients, s_val):
    """voxurf_coarse.py:384-433, use_mid=True.  sdf (N,S), gradients (N,S,3), steps (1,S) or (N,S)."""
    batch_size, n_samples = sdf.shape
    if steps.shape[0] == 1:
        steps = steps.repeat(batch_size, 1)
    dirs = viewdirs.unsqueeze(-2)
    inv_s = (torch.ones(1) / (torch.ones(1) * s_val)).expand(batch_size * n_samples, 1)
    true_cos = (dirs * gradients).sum(-1, keepdim=True)
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * 0.0 + F.relu(-true_cos) * 1.0).reshape(-1, 1)
    sdf = sdf.reshape(-1, 1)
    dists = steps[..., 1:] - steps[..., :-1]
    dists = torch.cat([dists, torch.Tensor([dists.mean()]).expand(dists[..., :1].shape)], -1)
    est_next = sdf + iter_cos * dists.reshape(-1, 1) * 0.5
    est_prev = sdf - iter_cos * dists.reshape(-1, 1) * 0.5
    prev_cdf = torch.sigmoid(est_prev * inv_s)
    next_cdf = torch.sigmoid(est_next * inv_s)
    p = prev_cdf - next_cdf
    return ((p + 1e-5) / (prev_cdf + 1e-5)).reshape(batch_size, n_samples).clip(0.0, 1.0)


def get_ray_marching_ray(alpha):
    """dvgo_ori.py:478-485: exclusive cumprod of clamp_min(1-alpha,1e-10)."""
    p = 1 - alpha
    alphainv_cum = torch.cat([torch.ones_like(p[..., [0]]), p.clamp_min(1e-10).cumprod(-1)], -1)
    return alpha * alphainv_cum[..., :-1], alphainv_cum


def coarse_forward_dense(m, rays_o, rays_d, viewdirs, global_step, *, near, stepsize, bg):
    """BASELINE.json config 1: the coarse model evaluated with the reference's dense PyTorch operators
    (sample_ray_ori + neus_alpha_from_sdf + get_ray_marching_ray) instead of the CUDA extension."""
    N = len(rays_o)
    pts, mask_outbbox, step = sample_ray_ori(rays_o, rays_d, m['xyz_min'], m['xyz_max'], m['sdf'].shape[2:],
                                             near, 1e9, stepsize, m['voxel_size'])
    S = pts.shape[1]
    flat = pts.reshape(-1, 3)
    sdf_grid = conv3d_replicate(m['sdf'], m['smooth_kernel']) if m.get('smooth_kernel') is not None else m['sdf']
    sdf = grid_trilinear(sdf_grid, flat, m['xyz_min'], m['xyz_max']).reshape(N, S)
    gradient = grid_trilinear(sdf_gradient_grid(m['sdf'], m['voxel_size']), flat, m['xyz_min'], m['xyz_max']).reshape(N, S, 3)
    s_val = float(torch.ones(1) * s_val_schedule(global_step, m['s_ratio'], m['s_start']))
    alpha = neus_alpha_from_sdf_dense(viewdirs, step, sdf, gradient, s_val)
    alpha = alpha * (~mask_outbbox)
    weights, alphainv_cum = get_ray_marching_ray(alpha)
    k0 = grid_trilinear(m['k0'], flat, m['xyz_min'], m['xyz_max'])
    rays_xyz = (flat - m['xyz_min']) / (m['xyz_max'] - m['xyz_min'])
    xyz_emb = positional_encoding(rays_xyz, m['posfreq'])
    viewdirs_emb = positional_encoding(viewdirs, m['viewfreq'])
    g = gradient.reshape(-1, 3)
    normal = g / (g.norm(dim=-1, keepdim=True) + 1e-5)
    feat = torch.cat([k0, xyz_emb, viewdirs_emb[:, None, :].expand(N, S, -1).reshape(N * S, -1), normal], -1)
    rgb = torch.sigmoid(mlp(feat, m['rgbnet'])).reshape(N, S, 3)
    rgb_marched = (weights[..., None] * rgb).sum(-2) + alphainv_cum[..., [-1]] * bg
    return {'rgb_marched': rgb_marched.clamp(0, 1), 'weights': weights, 'alphainv_cum': alphainv_cum, 'raw_alpha': alpha}


# ------------------------------------------------------------------------------------------------
# losses, regularisers, optimizer
# ------------------------------------------------------------------------------------------------
def total_variation(v, mask=None):
    """voxurf_fine.py:956-969."""
    tv2 = (v[:, :, 1:, :, :] - v[:, :, :-1, :, :]).abs()
    tv3 = (v[:, :, :, 1:, :] - v[:, :, :, :-1, :]).abs()
    tv4 = (v[:, :, :, :, 1:] - v[:, :, :, :, :-1]).abs()
    if mask is not None:
        tv2 = tv2[mask[:, :, :-1] & mask[:, :, 1:]]
        tv3 = tv3[mask[:, :, :, :-1] & mask[:, :, :, 1:]]
        tv4 = tv4[mask[:, :, :, :, :-1] & mask[:, :, :, :, 1:]]
    return (tv2.mean() + tv3.mean() + tv4.mean()) / 3


def total_variation_coarse(v, mask):
    """voxurf_coarse.py:702-715 -- NOT the fine file's formula: the three masked L1 sums are added and divided by
    3 * mask.sum() (the number of masked VOXELS, times C when the mask is repeated over channels), not by the per-axis
    pair counts."""
    tv2 = (v[:, :, 1:, :, :] - v[:, :, :-1, :, :]).abs()
    tv3 = (v[:, :, :, 1:, :] - v[:, :, :, :-1, :]).abs()
    tv4 = (v[:, :, :, :, 1:] - v[:, :, :, :, :-1]).abs()
    tv2 = tv2[mask[:, :, :-1] & mask[:, :, 1:]]
    tv3 = tv3[mask[:, :, :, :-1] & mask[:, :, :, 1:]]
    tv4 = tv4[mask[:, :, :, :, :-1] & mask[:, :, :, :, 1:]]
    return (tv2.sum() + tv3.sum() + tv4.sum()) / 3 / mask.sum()


def smooth_grad_tv(gradient, nonempty_mask, smooth_grad_tv_w):
    """voxurf_fine.py:417-420: ((tv_smooth_conv(G).detach() - G)[mask x3] ** 2).mean() * w.
    gradient (1,3,X,Y,Z) (autograd-connected to the sdf grid), nonempty_mask (1,1,X,Y,Z) bool."""
    g = gradient.permute(1, 0, 2, 3, 4)
    err = conv3d_replicate(g, tv_smooth_kernel()).detach() - g
    err = err[nonempty_mask.repeat(3, 1, 1, 1, 1)] ** 2
    return err.mean() * smooth_grad_tv_w


def fine_loss(ret, target, *, weight_main=1.0, weight_entropy_last=0.001, weight_rgb0=0.5):
    """run.py:604-636 for the surf/fine stage (weight_rgbper = 0).  The entropy term indexes [..., -1] on a
    1-D tensor, i.e. it reads the LAST RAY only (run.py:608)."""
    loss = weight_main * F.mse_loss(ret['rgb_marched'], target)
    if weight_entropy_last > 0:
        pout = ret['alphainv_cum'][..., -1].clamp(1e-6, 1 - 1e-6)
        ent = -(pout * torch.log(pout) + (1 - pout) * torch.log(1 - pout)).mean()
        loss = loss + weight_entropy_last * ent
    if weight_rgb0 > 0 and 'rgb_marched0' in ret:
        loss = loss + F.mse_loss(ret['rgb_marched0'], target) * weight_rgb0
    return loss


def python_adam_step(param, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.99, eps=1e-8, per_lr=None):
    """lib/utils.py:154-199 (the optimizer the reference trainer actually uses; dense, eps outside the
    bias-corrected sqrt).  In place; `step` is the already-incremented state step."""
    bias_correction1 = 1 - beta1 ** step
    bias_correction2 = 1 - beta2 ** step
    exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    denom = (exp_avg_sq.sqrt() / math.sqrt(bias_correction2)).add_(eps)
    step_size = lr / bias_correction1
    if per_lr is not None:
        param.addcdiv_(exp_avg * per_lr, denom, value=-step_size)
    else:
        param.addcdiv_(exp_avg, denom, value=-step_size)


def nonempty_mask(mc, xyz_min, xyz_max, shape):
    """voxurf_fine.py:353-361: mask-cache query at every lattice node (torch.linspace lattice). -> (1,1,X,Y,Z) bool"""
    xyz = torch.stack(torch.meshgrid(
        torch.linspace(float(xyz_min[0]), float(xyz_max[0]), shape[0]),
        torch.linspace(float(xyz_min[1]), float(xyz_max[1]), shape[1]),
        torch.linspace(float(xyz_min[2]), float(xyz_max[2]), shape[2]), indexing='ij'), -1)
    out = mask_cache_forward(mc['density'], xyz.reshape(-1, 3), mc['xyz_min'], mc['xyz_max'], mc['act_shift'],
                             mc['voxel_size_ratio'], mc['thres'])
    return out.reshape(1, 1, *shape)


# ------------------------------------------------------------------------------------------------
# caller side: view rays and the in-mask-cache ray filter (pinned by tests/golden/rays.npz)
# ------------------------------------------------------------------------------------------------
def view_rays(H, W, K, c2w, *, ndc=False, inverse_y=False, flip_x=False, flip_y=False, mode='center', jitter=None):
    """get_rays + get_rays_of_a_view, voxurf_fine.py:1001-1029,1065-1070 (ndc_rays :1043-1062 with near = 1).
    K (3,3), c2w (>=3,4) fp32 tensors.  jitter = (ji, jj) replaces the two torch.rand_like draws of mode 'random'.
    -> rays_o, rays_d, viewdirs (H,W,3)"""
    cols = torch.arange(W, dtype=torch.float32).view(1, W).expand(H, W)
    rows = torch.arange(H, dtype=torch.float32).view(H, 1).expand(H, W)
    if mode == 'center':
        u, v = cols + 0.5, rows + 0.5
    elif mode == 'lefttop':
        u, v = cols, rows
    elif mode == 'random':
        u, v = cols + jitter[0], rows + jitter[1]
    else:
        raise NotImplementedError(mode)
    if flip_x:
        u = torch.flip(u, dims=(1,))
    if flip_y:
        v = torch.flip(v, dims=(0,))
    x = (u - float(K[0][2])) / float(K[0][0])
    y = (v - float(K[1][2])) / float(K[1][1])
    cam = torch.stack([x, y, torch.ones_like(x)] if inverse_y else [x, -y, -torch.ones_like(x)], -1)
    R = c2w[:3, :3].float()
    prod = cam.unsqueeze(-2) * R                                   # (H,W,3,3): cam[l] * R[k][l]
    rays_d = (prod[..., 0] + prod[..., 1]) + prod[..., 2]
    rays_o = c2w[:3, 3].float().expand(rays_d.shape)
    sq = rays_d * rays_d
    viewdirs = rays_d / torch.sqrt((sq[..., 0] + sq[..., 1]) + sq[..., 2]).unsqueeze(-1)
    if ndc:
        focal, near = float(K[0][0]), 1.0
        cw, ch = -1. / (W / (2. * focal)), -1. / (H / (2. * focal))
        t = -(near + rays_o[..., 2]) / rays_d[..., 2]
        o = rays_o + t.unsqueeze(-1) * rays_d
        new_o = torch.stack([cw * o[..., 0] / o[..., 2], ch * o[..., 1] / o[..., 2], 1. + 2. * near / o[..., 2]], -1)
        new_d = torch.stack([cw * (rays_d[..., 0] / rays_d[..., 2] - o[..., 0] / o[..., 2]),
                             ch * (rays_d[..., 1] / rays_d[..., 2] - o[..., 1] / o[..., 2]),
                             -2. * near / o[..., 2]], -1)
        rays_o, rays_d = new_o, new_d
    return rays_o, rays_d, viewdirs


def hit_coarse_geo(mc, rays_o, rays_d, xyz_min, xyz_max, near, stepsize, voxel_size):
    """voxurf_fine.py:579-591: rays with at least one in-bbox sample inside the mask cache. -> bool, rays_o.shape[:-1]"""
    shape = rays_o.shape[:-1]
    o, d = rays_o.reshape(-1, 3).contiguous(), rays_d.reshape(-1, 3).contiguous()
    pts, mask_outbbox, ray_id = K.sample_pts_on_rays(o, d, xyz_min, xyz_max, near, 1e9, float(stepsize * voxel_size))[:3]
    inb = ~mask_outbbox
    inside = mask_cache_forward(mc['density'], pts[inb], mc['xyz_min'], mc['xyz_max'], mc['act_shift'],
                                mc['voxel_size_ratio'], mc['thres'])
    hit = torch.zeros(o.shape[0], dtype=torch.bool)
    hit[ray_id[inb][inside]] = True
    return hit.reshape(shape)


def training_rays_in_maskcache(images, poses, HW, Ks, mc, xyz_min, xyz_max, voxel_size, *, near, stepsize, ndc=False,
                               inverse_y=False, flip_x=False, flip_y=False):
    """get_training_rays_in_maskcache_sampling, voxurf_fine.py:1127-1164 -> rgb, rays_o, rays_d, viewdirs (n,3), counts"""
    out = [[], [], [], []]
    counts = []
    for img, c2w, (H, W), Kc in zip(images, poses, HW, Ks):
        ro, rd, vd = view_rays(H, W, Kc, c2w, ndc=ndc, inverse_y=inverse_y, flip_x=flip_x, flip_y=flip_y)
        keep = hit_coarse_geo(mc, ro, rd, xyz_min, xyz_max, near, stepsize, voxel_size)
        for dst, src in zip(out, (img, ro, rd, vd)):
            dst.append(src[keep])
        counts.append(int(keep.sum()))
    return [torch.cat(x) for x in out] + [counts]


# ------------------------------------------------------------------------------------------------
# off-step callers: mesh field query and progressive grid growing
# ------------------------------------------------------------------------------------------------
def sdf_field(sdf, xyz_min, xyz_max, resolution, smooth=True, sigma=0.5):
    """The field part of extract_geometry: voxurf_fine.py:894-910 (k = 3 Gaussian-smoothed grid, query of -sdf) on the
    lattice of dvgo_ori.py:679-693 (torch.linspace per axis). -> u (res, res, res)"""
    grid = conv3d_replicate(sdf, gaussian_kernel3d(3, sigma)) if smooth else sdf
    axes = [torch.linspace(float(xyz_min[i]), float(xyz_max[i]), resolution) for i in range(3)]
    pts = torch.stack(torch.meshgrid(*axes, indexing='ij'), -1).reshape(-1, 3)
    return grid_trilinear(-grid, pts, xyz_min, xyz_max).reshape(resolution, resolution, resolution)


def sdf_gradient_field(sdf, xyz_min, xyz_max, voxel_size, resolution, smooth=True, sigma=0.5):
    """The 6-tap trilinear gradient of the (smoothed) sdf grid on the same lattice: voxurf_fine.py:502-534 with
    sample_grad=True, displace 1.0 (BASELINE config 5's 'SDF + gradient field'). -> sdf (res,res,res), grad (res,res,res,3)"""
    grid = conv3d_replicate(sdf, gaussian_kernel3d(3, sigma)) if smooth else sdf
    axes = [torch.linspace(float(xyz_min[i]), float(xyz_max[i]), resolution) for i in range(3)]
    pts = torch.stack(torch.meshgrid(*axes, indexing='ij'), -1).reshape(-1, 3)
    s, g, _ = fine_grid_sampler(pts, grid, xyz_min, xyz_max, voxel_size)
    return s.reshape(resolution, resolution, resolution), g.reshape(resolution, resolution, resolution, 3)


def mesh_color_forward(m, pts):
    """voxurf_fine.py:804-892: colour of mesh vertices -- the fine forward's feature build and both MLPs at given
    points with viewdirs = -normal (normal = gradient / (|gradient| + 1e-5)), no compositing. -> rgb (P,3)"""
    sdf_grid = conv3d_replicate(m['sdf'], m['smooth_kernel']) if m.get('smooth_kernel') is not None else m['sdf']
    sdf, gradient, feat = fine_grid_sampler(pts, sdf_grid, m['xyz_min'], m['xyz_max'], m['voxel_size'])
    viewdirs = -(gradient / (gradient.norm(dim=-1, keepdim=True) + 1e-5))
    k0 = dense_grid_forward(m['k0'], pts, m['xyz_min'], m['xyz_max'])
    disp = sorted(set(m['grad_feat']))
    all_feat, all_grad = sample_sdfs(pts, sdf_grid, disp, m['xyz_min'], m['xyz_max'], m['voxel_size'],
                                     use_grad_norm=m['use_grad_norm'])
    rays_xyz = (pts - m['xyz_min']) / (m['xyz_max'] - m['xyz_min'])
    rgb_feat = torch.cat([positional_encoding(rays_xyz, m['posfreq']), positional_encoding(viewdirs, m['viewfreq'])], -1)
    hier = ([sdf[:, None]] if m['center_sdf'] else []) + [all_feat, all_grad]
    rgb_logit = mlp(torch.cat([rgb_feat, *hier], dim=-1), m['rgbnet'])
    k_feat = torch.cat([k0, positional_encoding(rays_xyz, m['k_posfreq']), positional_encoding(viewdirs, m['k_viewfreq']),
                        gradient], -1)
    if m.get('k_center_sdf', False):
        k_feat = torch.cat([k_feat, sdf[:, None]], -1)
    if m.get('k_res', True):
        k_feat = torch.cat([k_feat, rgb_logit.detach()], dim=-1)
    return torch.sigmoid(rgb_logit.detach() + mlp(k_feat, m['k_rgbnet']))


def scale_volume(grid, new_world_size):
    """grid.py:60-65: trilinear resampling of a (1,C,X,Y,Z) grid to the new resolution, align_corners=True"""
    return F.interpolate(grid, size=tuple(int(w) for w in new_world_size), mode='trilinear', align_corners=True)
