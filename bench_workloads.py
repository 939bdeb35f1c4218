"""bench.py --workload {render,mesh,coarse}: BASELINE.json's other configurations on the same contract (one JSON line).

  render  config 4 -- forward-only rendering of 800x800 views in 8192-ray chunks (run.py:81-227, chunks at :123-126) with
          render_grad + render_depth; views are dealt round-robin to the ranks, no exchange; a step = one view per rank.
          value = rays/s (device-resident images), e2e = rays/s with the rgb / depth / normal images copied to pinned host
          memory after every view (what run.py does with .cpu().numpy()).
  mesh    config 5 -- the 512^3 lattice of extract_geometry (lib/voxurf_fine.py:894-910 + lib/dvgo_ori.py:679-693): k=3
          sigma=0.5 smoothed -sdf and its 6-tap gradient, X-slabs dealt to the ranks (halo local: grids are replicated);
          a step = the whole lattice.  value = lattice points/s.
  coarse  config 2 -- the 96^3 coarse stage step (lib/voxurf_coarse.py:513-619 + run.py:600-659, configs/dtu_e2e/coarse.py):
          per-iteration 5^3 smoothing, gradient-grid sampling, weights recomputed after the threshold, autograd-form TV.
          value = iterations/s.

The reference arm of these workloads is the CPU oracle on a bounded sample, like the train workload's.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from voxurf_b200 import synthetic as S  # noqa: E402


def _dist_setup(world, local_rank):
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    import torch.distributed as dist
    if world > 1 and not dist.is_initialized():
        dist.init_process_group('nccl', device_id=device)
    return device, dist


def _barrier(world, dist):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(vals, world, dist, device):
    t = torch.tensor(vals, device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


# ------------------------------------------------------------------------------------------------ mesh (config 5)
def mesh_main(args, rank, world, local_rank):
    import bench as B
    res = args.mesh_res
    config = {'workload': f'mesh field query: {res}^3 lattice of extract_geometry over a {args.grid}^3 SDF grid, k=3 sigma=0.5 smoothed -sdf '
                          f'+ 6-tap gradient per lattice point, X-slabs over the ranks', 'grid': args.grid, 'lattice': res,
              'l2_policy': f'outputs larger than L2 ({res}^3 x 16 B per step); the {args.grid}^3 grid (67 MB at 256^3) is meant to stay L2-resident',
              'parallelism': f'{world} X-slabs, no exchange'}
    metric = f'lattice points/sec, SDF + gradient field ({res}^3 lattice)'
    if args.impl == 'reference':
        if rank != 0:
            return
        from oracle import voxurf_ref as R
        torch.set_num_threads(os.cpu_count())
        G = args.grid
        sdf = torch.from_numpy(S.sphere_sdf(G))
        mn, mx = torch.tensor([-1., -1., -1.]), torch.tensor([1., 1., 1.])
        vs = ((mx - mn).prod() / G ** 3).pow(1 / 3)
        r_s = min(res, 96)          # bounded sample: a (r_s)^3 lattice (the work is proportional to the number of points)
        ts = []
        for it in range(max(1, min(args.steps, 3)) + 1):
            t0 = time.perf_counter()
            R.sdf_gradient_field(sdf, mn, mx, vs, r_s, smooth=True, sigma=0.5)
            if it:
                ts.append(time.perf_counter() - t0)
        v = r_s ** 3 / float(np.mean(ts))
        cb = {'value': v, 'unit': 'points/s', 'cores': os.cpu_count(), 'kind': 'port',
              'sample': f'{r_s}^3 of the {res}^3 lattice per step (smoothing conv + 7 grid_sample taps per point), {len(ts)} steps'}
        print(json.dumps({'impl': 'reference', 'metric': metric, 'value': v, 'unit': 'points/s', 'n_gpus': args.gpus, 'steps': args.steps,
                          'warmup': args.warmup, 'ms_per_step': 1e3 * res ** 3 / v, 'higher_is_better': True, 'scaling': 'strong',
                          'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config, 'cpu_baseline': cb,
                          'e2e': {'value': v, 'unit': 'points/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return
    device, dist = _dist_setup(world, local_rank)
    from voxurf_b200 import _lib, parallel
    model = B.build_model(args, device)
    x0, x1 = parallel.shard_range(res, rank, world)

    def step(host=None):
        grid = model.mesh_query_grid(True, 0.5)            # the smoothing pass is part of the query (every rank, halo local)
        u, g = model.query_sdf_field(res, x_range=(x0, x1), with_gradient=True, sdf_grid=grid)
        if host is not None:
            host[0].copy_(u, non_blocking=True); host[1].copy_(g, non_blocking=True)
        return u, g

    for _ in range(max(args.warmup, 3)):
        step()
    _barrier(world, dist)
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    _barrier(world, dist)
    launches = _lib.launch_count() - l0
    host = (torch.empty(x1 - x0, res, res, pin_memory=True), torch.empty(x1 - x0, res, res, 3, pin_memory=True))
    step(host)
    _barrier(world, dist)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        step(host)
    f1.record()
    _barrier(world, dist)
    ms, ms_e2e = _max_over_ranks([e0.elapsed_time(e1), f0.elapsed_time(f1)], world, dist, device)
    if rank == 0:
        n_pts = res ** 3
        peak, peak_src = B.measured_peak()
        out_bytes = 16 * n_pts / world           # u + 3 gradient components written per point, per rank
        t = ms / args.steps * 1e-3
        # one rank's launch: the field kernel writes 16 B per lattice point of its slab (the grid reads hit L2)
        line = {'metric': metric, 'value': n_pts / t, 'unit': 'points/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32',
                'data': 'synthetic', 'config': config, 'gpu_launches': int(launches),
                'e2e': {'value': n_pts / (ms_e2e / args.steps * 1e-3), 'unit': 'points/s', 'h2d_bytes_per_step': 3 * 4 * res,
                        'd2h_bytes_per_step': int(16 * n_pts / world), 'ms_per_step': ms_e2e / args.steps},
                'roofline': {'bound': 'hbm', 'kernel': 'k_sdf_lattice (+ separable k=3 smoothing of the grid)', 'achieved': out_bytes / t / 1e9,
                             'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s', 'frac': out_bytes / t / 1e9 / peak,
                             'traffic': B.measured_traffic('k_sdf_lattice'),
                             'note': 'algorithmic bytes = 16 B written per lattice point (56 corner reads per point come from the L2-resident grid)'}}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ render (config 4)
def render_main(args, rank, world, local_rank):
    import bench as B
    H = W = 800
    N = args.rays
    rk = dict(near=2.0, far=6.0, bg=1.0, stepsize=0.5)     # lib/load_data.py:55 near/far, white background (nerf_synthetic_e2e/fine.py:16)
    n_chunks = (H * W + N - 1) // N
    config = {'workload': f'render-only: {H}x{W} views in {n_chunks} chunks of {N} rays (run.py:123-126), fine {args.grid}^3 SDF + '
                          f'{args.k0_channels}-ch k0, render_grad + render_depth, white bg, near/far 2/6', 'grid': args.grid,
              'k0_channels': args.k0_channels, 'rays_per_chunk': N, 'l2_policy': 'grids (0.87 GB) larger than L2; the 67 MB sdf grid is meant to stay L2-resident',
              'parallelism': f'views round-robin over {world} ranks, no exchange'}
    metric = f'rays/sec, render-only ({H}x{W} views, 8192-ray chunks)'
    if args.impl == 'reference':
        if rank != 0:
            return
        from oracle import voxurf_ref as R
        torch.set_num_threads(os.cpu_count())
        m, _, _, _ = B.oracle_bench_model(args.grid, args.k0_channels, 0)
        for k in ('sdf', 'k0'):
            m[k].requires_grad_(False)
        o, d, v = (torch.from_numpy(x) for x in S.make_rays(2048, seed=5))
        ts = []
        with torch.no_grad():
            for it in range(max(1, min(args.steps, 3)) + 1):
                t0 = time.perf_counter()
                R.fine_forward(m, o, d, v, None, near=0.3, stepsize=0.5, bg=1.0, render_grad=True, render_depth=True)
                if it:
                    ts.append(time.perf_counter() - t0)
        val = 2048 / float(np.mean(ts))
        cb = {'value': val, 'unit': 'rays/s', 'cores': os.cpu_count(), 'kind': 'port', 'sample': f'2048 rays per step (forward only), {len(ts)} steps'}
        print(json.dumps({'impl': 'reference', 'metric': metric, 'value': val, 'unit': 'rays/s', 'n_gpus': args.gpus, 'steps': args.steps,
                          'warmup': args.warmup, 'ms_per_step': 1e3 * H * W / val, 'higher_is_better': True, 'scaling': 'weak',
                          'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config, 'cpu_baseline': cb,
                          'e2e': {'value': val, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return
    device, dist = _dist_setup(world, local_rank)
    from voxurf_b200 import _lib, rays as RY
    from voxurf_b200.fused import FusedFineStep
    model = B.build_model(args, device)
    fused = FusedFineStep(model, N, None, rk, use_graph=not args.no_graph)
    views = [S.make_view(100 + i, H, W, inverse_y=False, r_cam=4.0, fov_scale=0.9) for i in range(max(args.steps, 4) * world)]

    def view_rays(i):
        _, _, K, c2w = views[i]
        ro, rd, vd = RY.get_rays_of_a_view(H, W, torch.from_numpy(K).to(device), torch.from_numpy(c2w).to(device), ndc=False,
                                           inverse_y=False, flip_x=False, flip_y=False)
        return ro.reshape(-1, 3), rd.reshape(-1, 3), vd.reshape(-1, 3)

    # calibrate the MLP row capacity on the densest chunk of a view (image centre)
    ro, rd, vd = view_rays(rank)
    mid = (n_chunks // 2) * N
    fused.calibrate(ro[mid:mid + N].contiguous(), rd[mid:mid + N].contiguous(), vd[mid:mid + N].contiguous(), headroom=1.5)
    rgb = torch.empty(n_chunks * N, 3, device=device); depth = torch.empty(n_chunks * N, device=device); normal = torch.empty(n_chunks * N, 3, device=device)
    pad = n_chunks * N - H * W

    def render_view(i, host=None):
        ro, rd, vd = view_rays(i)
        if pad:     # last chunk: repeat the first rays (their results are dropped)
            ro, rd, vd = (torch.cat([t, t[:pad]]) for t in (ro, rd, vd))
        for c in range(n_chunks):
            sl = slice(c * N, (c + 1) * N)
            out = fused.render_chunk(ro[sl], rd[sl], vd[sl])
            rgb[sl].copy_(out['rgb_marched']); depth[sl].copy_(out['depth']); normal[sl].copy_(out['normal_marched'])
        if host is not None:
            host[0].copy_(rgb[:H * W], non_blocking=True); host[1].copy_(depth[:H * W], non_blocking=True); host[2].copy_(normal[:H * W], non_blocking=True)

    for w_ in range(max(1, min(args.warmup, 2))):
        render_view(rank + world * w_)
    fused.poll_overflow(force=True)
    _barrier(world, dist)
    l0, r0 = _lib.launch_count(), fused.launches_replayed
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s_ in range(args.steps):
        render_view(rank + world * s_)
    e1.record()
    _barrier(world, dist)
    launches = _lib.launch_count() - l0 + fused.launches_replayed - r0
    host = (torch.empty(H * W, 3, pin_memory=True), torch.empty(H * W, pin_memory=True), torch.empty(H * W, 3, pin_memory=True))
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for s_ in range(args.steps):
        render_view(rank + world * s_, host)
    f1.record()
    _barrier(world, dist)
    fused.poll_overflow(force=True)
    ms, ms_e2e = _max_over_ranks([e0.elapsed_time(e1), f0.elapsed_time(f1)], world, dist, device)
    if rank == 0:
        total = H * W * world * args.steps
        M0, M2, M4 = fused.counts()
        line = {'metric': metric, 'value': total / (ms * 1e-3), 'unit': 'rays/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': config, 'execution': {'cuda_graph': bool(fused.use_graph)}, 'gpu_launches': int(launches),
                'views_per_sec': world * args.steps / (ms * 1e-3),
                'e2e': {'value': total / (ms_e2e * 1e-3), 'unit': 'rays/s', 'h2d_bytes_per_step': 2 * (36 + 64), 'd2h_bytes_per_step': H * W * 28,
                        'ms_per_step': ms_e2e / args.steps},
                'last_chunk_counts': {'M0': M0, 'M2': M2, 'M4': M4}}
        print(json.dumps(line))
    if world > 1:
        fused.release_graphs()
        dist.barrier()
        dist.destroy_process_group()


def main(args, rank, world, local_rank):
    if args.workload == 'mesh':
        return mesh_main(args, rank, world, local_rank)
    if args.workload == 'render':
        return render_main(args, rank, world, local_rank)
    if args.workload == 'coarse':
        import bench_coarse
        return bench_coarse.main(args, rank, world, local_rank)
    raise SystemExit('unknown workload')
