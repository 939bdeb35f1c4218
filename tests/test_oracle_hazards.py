"""CPU: hand-derived expectations for the parity hazards of SURVEY Appendix A on the C oracle (oracle/ref_kernels.c).
The golden vectors pin the oracle to the reference's outputs; these pin it to the reference's documented *semantics* on
crafted inputs, so that a product kernel that matches the oracle (tests/test_gpu_*.py) matches them too."""
import numpy as np
import torch

from oracle import kernels as K

XYZ_MIN, XYZ_MAX = torch.tensor([-1., -1., -1.]), torch.tensor([1., 1., 1.])


def test_hazard_1_2_3_4_sampling():
    o = torch.tensor([[0., 0., -3.], [5., 5., 5.], [-3., 1., 0.], [0., 0., -3.]])
    d = torch.tensor([[0., 0., 1.], [1., 0.2, 0.1], [1., 0., 0.], [0., 0., 2.]])
    t_min, t_max = K.infer_t_minmax(o, d, XYZ_MIN, XYZ_MAX, 0.2, 1e9)
    # hazard 1: zero direction components are replaced by 1e-6: the slab test of the axes the ray does not move along
    # gives +-2e6 / +-1e6 and never decides; ray 0 enters at z = -1 (t = 2) and leaves at z = +1 (t = 4)
    assert float(t_min[0]) == 2.0 and float(t_max[0]) == 4.0
    # near is honoured (hazard 4): a ray starting inside would get t_min = near
    tn, _ = K.infer_t_minmax(torch.zeros(1, 3), torch.tensor([[0., 0., 1.]]), XYZ_MIN, XYZ_MAX, 0.2, 1e9)
    assert abs(float(tn[0]) - 0.2) < 1e-7
    stepdist = 0.05
    pts, mask_out, ray_id, step_id, n_steps, _, _ = K.sample_pts_on_rays(o, d, XYZ_MIN, XYZ_MAX, 0.2, 1e9, stepdist)
    # hazard 2: the ray that misses the box still gets exactly one sample, flagged out of the box
    assert int(n_steps[1]) == 1 and bool(mask_out[ray_id == 1].all()) and int((ray_id == 1).sum()) == 1
    # hazard 3: strict comparisons -- ray 2 runs inside the face y = +1 and all its samples within x in [-1, 1] are inside
    sel = ray_id == 2
    inside = (pts[sel, 0] >= -1) & (pts[sel, 0] <= 1)
    assert bool((~mask_out[sel])[inside].all())
    # N_steps = max(ceil((t_max - t_min) * |d| / stepdist), 1); direction length matters (ray 3 = ray 0 with |d| = 2)
    assert int(n_steps[0]) == int(np.ceil((4.0 - 2.0) * 1.0 / stepdist)) and int(n_steps[3]) == int(np.ceil((2.0 - 1.0) * 2.0 / stepdist))
    assert ray_id.dtype == torch.int64 and step_id.dtype == torch.int64 and mask_out.dtype == torch.bool


def test_hazard_5_maskcache_lookup_rounding_and_range():
    world = torch.zeros(3, 3, 3, dtype=torch.bool)
    world[2, 1, 0] = True
    scale, shift = torch.tensor([1., 1., 1.]), torch.tensor([0., 0., 0.])     # ijk = round(xyz * 1 + 0)
    q = torch.tensor([[1.5, 0.5, -0.5],      # round half away from zero: (2, 1, -1) -> out of range -> False
                      [1.5, 0.5, 0.49],      # (2, 1, 0) -> True
                      [2.49, 1.49, -0.49],   # (2, 1, 0) (-0.49 rounds to 0) -> True
                      [2.5, 1.0, 0.0],       # (3, 1, 0) -> out of range -> False
                      [-0.5, 1.0, 0.0]])     # (-1, ...) -> False
    out = K.maskcache_lookup(world, q, scale, shift)
    assert out.tolist() == [False, True, True, False, False]


def test_hazard_6_alpha2weight_early_exit_and_tail():
    # one ray of 6 samples: alpha = 1 at index 2 drives T to 0 -> the loop stops after writing that sample
    alpha = torch.tensor([0.5, 0.5, 1.0, 0.3, 0.3, 0.3, 0.25, 0.0])
    ray_id = torch.tensor([0, 0, 0, 0, 0, 0, 1, 1])
    w, T, last, i_start, i_end = K.alpha2weight(alpha, ray_id, 3)
    assert w[:3].tolist() == [0.5, 0.25, 0.25] and T[:3].tolist() == [1.0, 0.5, 0.25]
    assert w[3:6].tolist() == [0.0, 0.0, 0.0] and T[3:6].tolist() == [1.0, 1.0, 1.0]      # tail: w = 0, T = 1
    assert int(i_start[0]) == 0 and int(i_end[0]) == 3                                     # i_end moved to the exit
    assert float(last[0]) == 0.0                                                           # T at exit
    assert w[6:].tolist() == [0.25, 0.0] and float(last[1]) == 0.75 and int(i_end[1]) == 8
    assert float(last[2]) == 1.0 and int(i_start[2]) == int(i_end[2])                      # empty ray: nothing absorbed


def test_hazard_7_alpha2weight_backward_denominator():
    alpha = torch.tensor([0.5, 1.0])       # 1 - alpha + 1e-10 keeps the division finite at alpha = 1 (double arithmetic)
    ray_id = torch.tensor([0, 0])
    fw = K.alpha2weight(alpha, ray_id, 1)
    g = K.alpha2weight_backward(alpha, *fw, 1, torch.tensor([1.0, 1.0]), torch.tensor([0.5]))
    assert torch.isfinite(g).all()


def test_hazard_17_tv_axis_weight_quirk():
    # unmasked kernel: the fastest and the slowest axis both use wz, wx is ignored; weights are divided by 6
    p = torch.zeros(1, 1, 3, 3, 3)
    p[0, 0, 1, 1, 1] = 0.6
    base = torch.zeros_like(p)
    g1 = base.clone(); K.total_variation_add_grad(p, g1, 100.0, 0.0, 0.0, True)
    assert float(g1.abs().max()) == 0.0                                   # wx alone does nothing
    g2 = base.clone(); K.total_variation_add_grad(p, g2, 0.0, 0.0, 6.0, True)
    # centre voxel: 2 neighbours on the fastest axis + 2 on the slowest, each clamp(0.6, -1, 1) * wz / 6
    assert abs(float(g2[0, 0, 1, 1, 1]) - 4 * 0.6) < 1e-6
    g3 = base.clone(); K.total_variation_add_grad(p, g3, 0.0, 6.0, 0.0, True)
    assert abs(float(g3[0, 0, 1, 1, 1]) - 2 * 0.6) < 1e-6                 # wy: the middle axis only


def test_hazard_18_adam_epsilon_placement():
    # CUDA kernel (adam_upd_kernel.cu:21,72): p -= step_size * m / (sqrt(v) + eps), step_size = lr * sqrt(1-b2^t) / (1-b1^t)
    p, g = torch.tensor([1.0]), torch.tensor([1e-6])
    m, v = torch.zeros(1), torch.zeros(1)
    K.adam_upd(p, g, m, v, 1, 0.9, 0.99, 0.1, 1e-8, mode=0)
    step_size = 0.1 * np.sqrt(1 - 0.99) / (1 - 0.9)
    expect = 1.0 - step_size * (0.1 * 1e-6) / (np.sqrt(0.01 * 1e-12) + 1e-8)
    assert abs(float(p[0]) - expect) < 1e-6
    # the Python optimizer (utils.py:190-199) divides sqrt(v) by sqrt(bias_correction2) *before* adding eps: different for tiny g
    from oracle import voxurf_ref as R
    p2, m2, v2 = torch.tensor([1.0]), torch.zeros(1), torch.zeros(1)
    R.python_adam_step(p2, g.clone(), m2, v2, 1, 0.1)
    expect2 = 1.0 - (0.1 / (1 - 0.9)) * (0.1 * 1e-6) / (np.sqrt(0.01 * 1e-12) / np.sqrt(1 - 0.99) + 1e-8)
    assert abs(float(p2[0]) - expect2) < 1e-6 and abs(expect - expect2) > 1e-4


# ---------------------------------------------------------------------------------------------- torch half of the oracle
def test_hazard_10_12_fd_gradient_boundaries_and_replicate_padding():
    from oracle import voxurf_ref as R
    rs = np.random.RandomState(0)
    sdf = torch.from_numpy(rs.standard_normal((1, 1, 5, 6, 7)).astype(np.float32))
    g = R.sdf_gradient_grid(sdf, torch.tensor(0.25))
    # hazard 10: each component is zero on the two boundary planes of its own axis, central difference / 2 / voxel elsewhere
    assert float(g[0, 0, 0].abs().max()) == 0 and float(g[0, 0, -1].abs().max()) == 0
    assert float(g[0, 1, :, 0].abs().max()) == 0 and float(g[0, 1, :, -1].abs().max()) == 0
    assert float(g[0, 2, :, :, 0].abs().max()) == 0 and float(g[0, 2, :, :, -1].abs().max()) == 0
    assert abs(float(g[0, 0, 2, 3, 4]) - float((sdf[0, 0, 3, 3, 4] - sdf[0, 0, 1, 3, 4]) / 2 / 0.25)) < 1e-6
    # hazard 12: replicate padding -- a constant grid stays constant under the normalised smoothing kernel, borders included
    c = torch.full((1, 1, 4, 5, 6), 0.7)
    out = R.conv3d_replicate(c, R.gaussian_kernel3d(5, 0.8))
    assert out.shape == c.shape and float((out - 0.7).abs().max()) < 1e-6


def test_hazard_14_16_alpha_epsilons_and_last_ray_entropy():
    from oracle import voxurf_ref as R
    # hazard 14: (prev - next + 1e-5) / (prev + 1e-5), clipped to [0, 1]; a sample the ray leaves the surface through
    # (gradient along the view direction) has iter_cos = 0 -> prev == next -> alpha = 1e-5 / (prev + 1e-5)
    vd = torch.tensor([[0., 0., 1.]])
    a = R.neus_alpha_from_sdf_scatter(vd, torch.tensor([0]), 0.1, torch.tensor([0.0]), torch.tensor([[0., 0., 1.]]), 0.05)
    assert abs(float(a[0]) - 1e-5 / (0.5 + 1e-5)) < 1e-9
    a = R.neus_alpha_from_sdf_scatter(vd, torch.tensor([0]), 0.1, torch.tensor([0.0]), torch.tensor([[0., 0., -1.]]), 0.05)
    prev, nxt = 1 / (1 + np.exp(-0.05 / 0.05)), 1 / (1 + np.exp(0.05 / 0.05))
    assert abs(float(a[0]) - (prev - nxt + 1e-5) / (prev + 1e-5)) < 1e-6
    # hazard 16: the entropy term reads the last ray only
    target = torch.zeros(3, 3)
    ret = {'rgb_marched': torch.zeros(3, 3), 'rgb_marched0': torch.zeros(3, 3), 'alphainv_cum': torch.tensor([0.3, 0.6, 0.5])}
    base = float(R.fine_loss(ret, target))
    ret['alphainv_cum'] = torch.tensor([0.9, 0.1, 0.5])
    assert float(R.fine_loss(ret, target)) == base
    ret['alphainv_cum'] = torch.tensor([0.3, 0.6, 0.9])
    assert float(R.fine_loss(ret, target)) != base
    assert abs(base - 0.001 * np.log(2.0)) < 1e-7          # entropy of p = 0.5, weight 0.001
