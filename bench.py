#!/usr/bin/env python
"""bench.py -- Voxurf fine-stage training step (fwd + bwd + TV + per-voxel Adam) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): rays/sec for one 8192-ray batch per GPU per step, fine stage, 256^3 grids.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.

Arms
  ours       the CUDA path of this repo through its public API (voxurf_b200.trainer.Trainer.step).
  reference  the reference's own formulation of the same step restated for CPU (oracle/voxurf_ref.py +
             oracle/ref_kernels.c), timed on the host cores of this box on a bounded ray sample.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from voxurf_b200 import synthetic as S  # noqa: E402

G_FINE, N_RAYS, WIDTH = 256, 8192, 192
START_STEP = 15001   # right after the 160^3 -> 256^3 growth (configs/dtu_e2e/fine.py:26-27)
RENDER_KW = dict(near=0.3, far=6.0, bg=0.0, stepsize=0.5)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=6)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--k0-channels', type=int, default=12, help='12 = BASELINE.json / default_fine_s.py:111; 6 = dtu_e2e/fine.py:73')
    ap.add_argument('--grid', type=int, default=G_FINE)
    ap.add_argument('--rays', type=int, default=N_RAYS)
    ap.add_argument('--smooth', type=int, default=0, help='per-iteration Gaussian smoothing ksize (0 = reference fine config)')
    ap.add_argument('--layout', default='channels_last', choices=['channels_last', 'channel_major'])
    ap.add_argument('--cpu-rays', type=int, default=256, help='ray sample of the CPU baseline')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--path', default='fused', choices=['fused', 'dropin'], help='fused = sync-free FusedFineStep; dropin = Voxurf.forward + autograd')
    ap.add_argument('--graph-multi', action='store_true', help='N > 1: capture the step including its NCCL collectives (experimental)')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel of the step separately (no CUDA-graph replay)')
    ap.add_argument('--dense-adam', action='store_true', help='k0 Adam over every voxel (no touched/live bitmaps)')
    ap.add_argument('--dense-k0-allreduce', action='store_true', help='multi-GPU: all-reduce the dense k0 gradient grid instead of exchanging rows')
    ap.add_argument('--phases', action='store_true', help='also print a per-phase CUDA-event breakdown to stderr')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ scene
def mask_state():
    d = torch.from_numpy(S.mask_density(100))
    return {'MaskCache_kwargs': {'xyz_min': [-1., -1., -1.], 'xyz_max': [1., 1., 1.], 'act_shift': float(np.log(1 / (1 - 1e-6) - 1)),
                                 'voxel_size_ratio': 1.0, 'nearest': False}, 'model_state_dict': {'density': d}}


def build_model(args, device):
    from voxurf_b200 import voxurf_fine as VF
    cfg = {k: v for k, v in S.FINE_CFG.items() if k != 'stepsize'}
    G = args.grid
    torch.manual_seed(0)
    m = VF.Voxurf(xyz_min=[-1., -1., -1.], xyz_max=[1., 1., 1.], num_voxels=G ** 3, num_voxels_base=G ** 3,
                  rgbnet_dim=args.k0_channels, rgbnet_width=WIDTH, smooth_ksize=args.smooth, smooth_sigma=0.8,
                  k0_channels_last=(args.layout == 'channels_last'), mask_cache_state=mask_state(), **cfg)
    assert tuple(int(w) for w in m.world_size) == (G, G, G), m.world_size
    m = m.to(device)
    with torch.no_grad():
        m.sdf.grid.data = (m.sdf.grid.data + 1.0 - 0.5) / 0.3   # ball init is r - 1; scene is (r - 0.5) / sdf_reduce
        gen = torch.Generator(device=device).manual_seed(1234)
        m.k0.grid.data.copy_(0.1 * torch.randn(m.k0.grid.shape, generator=gen, device=device))
        m._set_nonempty_mask()
    return m


def ray_pool(n_batches, n_rays, rank):
    pool = []
    for b in range(n_batches):
        o, d, v = S.make_rays(n_rays, seed=777 + 1000 * rank + b)
        pool.append(tuple(torch.from_numpy(x) for x in (o, d, v, S.make_target(v, seed=b))))
    return pool


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], r[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------ roofline helpers
def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


def measured_tensor_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['bf16_tflops_sustained']), 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)'
        except Exception:
            pass
    return 1590.0, 'fallback (B200_PROFILING.md, 1.59 PFLOP/s bf16)'


def algorithmic_bytes(V, C, tv, M2, M3, M4, N, fd):
    """B = 4V[9(1+C) + 3 tv + 4 fd] + (16 M2 + 8 M3 + 20 M4) + 60 N"""
    return 4 * V * (9 * (1 + C) + 3 * tv + 4 * fd) + (16 * M2 + 8 * M3 + 20 * M4) + 60 * N


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_arm(args, steps, warmup, n_rays_sample):
    """The reference's formulation of the step restated for CPU (oracle), on a bounded sample.

    The step has a grid-proportional part (full-grid FD gradient, TV, dense Adam over sdf + k0: independent of the
    batch) and a ray-proportional part.  We time one full grid part and the ray part on `n_rays_sample` rays, and
    report rays/s for the full 8192-ray batch:  8192 / (t_grid + (8192 / n_rays_sample) * t_rays)."""
    from oracle import kernels as K
    from oracle import voxurf_ref as R
    import torch.nn.functional as F
    torch.set_num_threads(os.cpu_count())
    G, C = args.grid, args.k0_channels
    xyz_min, xyz_max = torch.tensor([-1., -1., -1.]), torch.tensor([1., 1., 1.])
    cfg = S.FINE_CFG
    rs = np.random.RandomState(0)
    dim0, k_dim0 = S.fine_dims(C)
    lay = lambda ls: [(torch.from_numpy(W).requires_grad_(True), torch.from_numpy(b).requires_grad_(True)) for W, b in ls]
    mc = dict(density=R.mask_cache_density(torch.from_numpy(S.mask_density(100))), xyz_min=xyz_min, xyz_max=xyz_max,
              act_shift=float(np.log(1 / (1 - 1e-6) - 1)), voxel_size_ratio=1.0, thres=1e-3)
    voxel_size = ((xyz_max - xyz_min).prod() / (G ** 3)).pow(1 / 3)
    sdf = torch.from_numpy(S.sphere_sdf(G))
    nonempty = R.nonempty_mask(mc, xyz_min, xyz_max, (G, G, G))
    sdf[~nonempty] = 1
    m = dict(xyz_min=xyz_min, xyz_max=xyz_max, voxel_size=voxel_size, mask_cache=mc, nonempty_mask=nonempty,
             sdf=sdf.requires_grad_(True), k0=(0.1 * torch.randn(1, C, G, G, G)).requires_grad_(True),
             rgbnet=lay(S.mlp_init(rs, dim0, WIDTH, 4)), k_rgbnet=lay(S.mlp_init(rs, k_dim0, WIDTH, 4)),
             posfreq=torch.FloatTensor([2 ** i for i in range(5)]), viewfreq=torch.FloatTensor([1.]),
             k_posfreq=torch.FloatTensor([2 ** i for i in range(5)]), k_viewfreq=torch.FloatTensor([1.]),
             grad_feat=cfg['grad_feat'], use_grad_norm=True, center_sdf=True, k_center_sdf=False, k_res=True,
             fast_color_thres=1e-4, s_ratio=50, s_start=0.05, step_start=0, s_val=0.05,
             smooth_kernel=R.gaussian_kernel3d(args.smooth, 0.8) if args.smooth > 0 else None)
    params = [m['sdf'], m['k0']] + [t for W, b in m['rgbnet'] + m['k_rgbnet'] for t in (W, b)]
    lrs = [5e-3, 1e-1] + [3e-3] * 8 + [1e-3] * 8
    state = [(torch.zeros_like(p), torch.zeros_like(p)) for p in params]
    o, d, v = (torch.from_numpy(x) for x in S.make_rays(n_rays_sample, seed=777))
    target = torch.from_numpy(S.make_target(v.numpy()))
    t_rays, t_grid = [], []
    for it in range(warmup + steps):
        gs = START_STEP + it
        for p in params:
            p.grad = None
        t0 = time.perf_counter()
        ret = R.fine_forward(m, o, d, v, gs, near=RENDER_KW['near'], stepsize=0.5, bg=0.0)
        loss = R.fine_loss(ret, target)
        loss.backward()
        t1 = time.perf_counter()
        # grid-proportional: smooth-grad TV (through the full FD gradient), TV add-grad, dense Adam
        tvl = 0.01 * R.smooth_grad_tv(R.sdf_gradient_grid(m['sdf'], voxel_size), nonempty, 0.05)
        tvl.backward()
        w = 0.01 * 0.1 / N_RAYS * G / 128
        K.total_variation_add_grad(m['sdf'].detach(), m['sdf'].grad, w, w, w, True)
        with torch.no_grad():
            for p, lr, (ea, es) in zip(params, lrs, state):
                if p.grad is not None:
                    R.python_adam_step(p, p.grad, ea, es, it + 1, lr)
        t2 = time.perf_counter()
        if it >= warmup:
            t_rays.append(t1 - t0); t_grid.append(t2 - t1)
    tr, tg = float(np.mean(t_rays)), float(np.mean(t_grid))
    full = tg + (N_RAYS / n_rays_sample) * tr
    return {'value': N_RAYS / full, 'unit': 'rays/s', 'cores': os.cpu_count(), 'kind': 'port',
            'sample': f'{n_rays_sample} of {N_RAYS} rays per step ({tr:.2f} s, incl. dense 256^3 autograd grads) + one full '
                      f'grid-proportional part ({tg:.2f} s: FD gradient + smooth-grad TV + TV add-grad + dense Adam); '
                      f'extrapolated to the 8192-ray batch; {steps} steps after {warmup} warm-up, torch {torch.get_num_threads()} threads',
            't_rays_s': tr, 't_grid_s': tg}


# ------------------------------------------------------------------------------------------------ main
def main():
    args = parse()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    C, G = args.k0_channels, args.grid
    workload = (f'voxurf_fine fwd+bwd+TV+Adam, {G}^3 SDF + {C}-ch k0, rgbnet 79->192^3->3 + k_rgbnet, {args.rays}-ray batch/GPU, '
                f'synthetic sphere scene, step {START_STEP}+, TV every 3rd iter, smooth_ksize={args.smooth}')
    config = {'workload': workload, 'path': args.path, 'grid': G, 'k0_channels': C, 'rays_per_gpu': args.rays, 'k0_layout': args.layout,
              'l2_policy': 'inputs larger than L2 (grids + moments 0.27 GB x (1+C) >> 126 MB)', 'parallelism': f'dp{world} over rays'}

    if args.impl == 'reference':
        if rank != 0:
            return
        # every step is a bounded sample (256 rays + one full grid-proportional part, ~1.4 s of CPU work): K steps are
        # honoured up to 30 so that the run stays within a few minutes
        k_used, w_used = max(1, min(args.steps, 30)), max(0, min(args.warmup, 3))
        cb = cpu_reference_arm(args, k_used, w_used, args.cpu_rays)
        line = {'impl': 'reference', 'metric': 'rays/sec fwd+bwd+Adam (fine 256^3, 8192-ray batch)', 'value': cb['value'], 'unit': 'rays/s',
                'n_gpus': args.gpus, 'steps': k_used, 'warmup': w_used, 'ms_per_step': 1e3 * N_RAYS / cb['value'],
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
                'cpu_baseline': cb, 'e2e': {'value': cb['value'], 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl ours needs a CUDA device (there is no CPU fallback)')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    from voxurf_b200 import _lib
    from voxurf_b200.trainer import FINE_TRAIN, Trainer
    from voxurf_b200 import parallel

    model = build_model(args, device)
    pool = ray_pool(8, args.rays, rank)
    dev_pool = [tuple(t.to(device) for t in b) for b in pool]
    pin_pool = [tuple(t.pin_memory() for t in b) for b in pool]
    sync = parallel.GradSync(model, world) if world > 1 else None
    fused = None
    if args.path == 'fused':
        from voxurf_b200.fused import FusedFineStep
        fused = FusedFineStep(model, args.rays, FINE_TRAIN, RENDER_KW, world=world, rank=rank, sparse_adam=not args.dense_adam,
                              use_graph=((world == 1 or args.graph_multi) and not args.no_graph), graph_multi_gpu=args.graph_multi)
        fused.calibrate(*dev_pool[0][:3], global_step=START_STEP, headroom=1.35)
        sync = None   # FusedFineStep.grad_sync(): dense all-reduce for sdf + MLPs, row exchange for k0 (or --dense-k0-allreduce)
        fused.sparse_k0_exchange = not args.dense_k0_allreduce

        class _T:   # same surface as Trainer for the loops below
            optimizer = fused

            @staticmethod
            def step(ro, rd, vd, tg, global_step):
                return fused.step(ro, rd, vd, tg, global_step, grad_sync=sync), None
        trainer = _T
    else:
        trainer = Trainer(model, FINE_TRAIN, RENDER_KW, zero_grad_in_step=False, grad_sync=sync)
        trainer.global_batch = args.rays * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(n, first_step, e2e):
        loss_host = torch.empty(1, pin_memory=True)
        for i in range(n):
            b = (first_step + i) % len(pool)
            if e2e:
                batch = tuple(t.to(device, non_blocking=True) for t in pin_pool[b])
            else:
                batch = dev_pool[b]
            loss, ret = trainer.step(*batch, global_step=first_step + i)
            if e2e:
                loss_host.copy_(loss.reshape(1), non_blocking=False)
        return ret

    # ---- device-resident timing
    run(args.warmup, START_STEP, False)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.launch_count()
    from voxurf_b200 import mlp as _mlp
    graph = fused is not None and fused.use_graph
    config['cuda_graph'] = bool(graph)   # the timed steps are replays of two captured graphs (TV / non-TV iteration)
    r0 = fused.launches_replayed if graph else 0
    if fused is None:
        trainer.optimizer.timed_param = model.k0.grid
    if not graph:                       # per-kernel event timing inside the timed steps
        trainer.optimizer.timings = []
        if fused is not None:
            _mlp.CHAIN_TIMINGS = []
    with ClockSampler(local_rank) as clk:
        ev0.record()
        ret = run(args.steps, START_STEP + args.warmup, False)
        ev1.record()
        barrier()
    launches = _lib.launch_count() - l0 + ((fused.launches_replayed - r0) if graph else 0)
    ms = ev0.elapsed_time(ev1)
    if graph:
        # the timed steps were CUDA-graph replays (no events inside a graph): time the same kernels with events over the
        # same number of separately launched steps right after, outside the headline measurement
        fused.timings, _mlp.CHAIN_TIMINGS = [], []
        run(args.steps, START_STEP + args.warmup + 3 * args.steps, False)
        barrier()
    tm = trainer.optimizer.timings
    if fused is not None:
        adam_ms = [ev[0].elapsed_time(ev[1]) for name, ev in tm if name == 'k0']
        sdf_adam_ms = [ev[0].elapsed_time(ev[1]) for name, ev in tm if name == 'sdf']
    else:
        adam_ms, sdf_adam_ms = [a.elapsed_time(b) for a, b in tm], []
    chain = [((a.elapsed_time(b)), fl) for (a, b), fl in (_mlp.CHAIN_TIMINGS or [])]
    _mlp.CHAIN_TIMINGS = None
    if fused is None:
        trainer.optimizer.timed_param = None
    else:
        fused.timings = None
    # ---- end-to-end timing (pinned host -> device inputs, device -> host loss, every step)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(args.steps, START_STEP + args.warmup + args.steps, True)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    probe = None
    if fused is not None and fused.k0_touched is not None:
        # one extra (untimed) step on every rank to count the voxels the sparse-aware k0 Adam pass touches
        fused.bitmap_probe = []
        run(1, START_STEP + args.warmup + 2 * args.steps, False)
        torch.cuda.synchronize()
        probe, fused.bitmap_probe = fused.bitmap_probe, None
    if rank == 0:
        total_rays = args.rays * world * args.steps
        value = total_rays / (ms * 1e-3)
        V = G ** 3
        peak, peak_src = measured_peak()
        # SURVEY.md 8d counts 7 Adam passes (read p,g,m,v; write p,m,v) + 1 gradient zero-fill pass per element; this kernel
        # does all 8 in one pass -> 32 B/element.  profiles/r01c: dram traffic 6.383 GB per launch vs 6.442 GB algorithmic.
        adam_bytes = 32 * V * C
        adam_kernel = 'k_adam (k0 grid, %d x %d^3 fp32; Adam + fused grad zero-fill, 32 B/element)' % (C, G)
        sparse = None
        if fused is not None and fused.k0_touched is not None:
            # sparse-aware pass: 32 B/element where a gradient landed this step, 24 B/element for voxels that ever had
            # one (non-zero moments), nothing elsewhere (identity update), plus both bitmaps.  Counted on one extra step.
            tb, lb = (np.unpackbits(x.cpu().numpy().view(np.uint8)) for x in probe[0])
            n_t, n_l = int(tb.sum()), int((tb | lb).sum())
            adam_bytes = 4 * C * (8 * n_t + 6 * (n_l - n_t)) + 2 * 4 * fused.k0_touched.numel()
            adam_kernel = 'k_adam (k0 grid, %d x %d^3 fp32, sparse-aware: %d touched voxels x 32 B/el + %d live x 24 B/el + bitmaps)' % (C, G, n_t, n_l - n_t)
            sparse = {'voxels': V, 'touched': n_t, 'live': n_l, 'dense_equivalent_bytes': 32 * V * C}
        adam_t = float(np.mean(adam_ms)) if adam_ms else None
        step_ms = ms / args.steps
        roof_k0 = {'bound': 'hbm', 'kernel': adam_kernel, 'achieved': (adam_bytes / (adam_t * 1e-3) / 1e9) if adam_t else None,
                   'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s', 'traffic': (6.383e9 if (C == 12 and G == 256 and sparse is None) else None),
                   'ms_per_launch': adam_t, 'algorithmic_bytes_per_launch': adam_bytes, 'share_of_step': (adam_t / step_ms) if adam_t else None}
        roof_k0['frac'] = roof_k0['achieved'] / peak if roof_k0['achieved'] else None
        if sparse:
            roof_k0['sparse'] = sparse
        if fused is not None:
            M0, M2, M4 = fused.counts()
            M3 = int(fused.keep[:M2].sum())
        else:
            M4 = int(ret['weights'].shape[0]); M3 = int(ret['mask'].shape[0]); M0 = int(ret['mask_outbbox'].shape[0])
            M2 = int((~ret['mask_outbbox']).sum())
        roof, extra = roof_k0, {}
        if chain and sum(t for t, _ in chain) / args.steps > (adam_t or 0):
            # the dominant kernel of the fused step is the tcgen05 layer-chain kernel (4 launches per step: two forward
            # chains, two dX chains).  Algorithmic flops = 2 * rows * sum_l K_l N_l in fp32 terms; the kernel issues
            # three TF32 MMAs per product term (TF32x3 split), and TF32 runs at half the bf16 rate, so the ceiling of
            # this formulation is peak / 6.  peak = the driver-measured dense bf16 cuBLAS throughput (sustained figure:
            # the kernel is timed inside a long step).
            tf_peak, tf_src = measured_tensor_peak()
            t_chain = float(np.mean([t for t, _ in chain]))
            fl = float(np.mean([f for _, f in chain])) * M4
            roof = {'bound': 'tensor', 'kernel': 'k_mlp_chain (fused 4-layer MLP chain on tcgen05, TF32x3; %d rows, 4 launches/step)' % M4,
                    'achieved': fl / (t_chain * 1e-3) / 1e12, 'peak': tf_peak, 'peak_source': tf_src, 'unit': 'TFLOP/s',
                    # dram read + write bytes per launch, mean of the four chain launches of one step, from
                    # profiles/r01j_final_launches.md (ncu --set full); only valid for the profiled configuration
                    'traffic': (1.693e8 if (C == 12 and G == 256 and args.rays == 8192) else None), 'ms_per_launch': t_chain,
                    'algorithmic_flops_per_launch': fl,
                    'issued_tf32_flops_per_launch': 3 * fl, 'formulation_ceiling_frac': 1.0 / 6.0,
                    'share_of_step': sum(t for t, _ in chain) / args.steps / step_ms}
            roof['frac'] = roof['achieved'] / tf_peak
            extra['roofline_k0_adam'] = roof_k0
            if sdf_adam_ms:
                t_sdf = float(np.mean(sdf_adam_ms))
                extra['roofline_sdf_adam'] = {'bound': 'hbm', 'kernel': 'k_adam (sdf grid, %d^3 fp32, dense, 32 B/element)' % G,
                                              'achieved': 32 * V / (t_sdf * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                                              'frac': 32 * V / (t_sdf * 1e-3) / 1e9 / peak, 'ms_per_launch': t_sdf,
                                              'algorithmic_bytes_per_launch': 32 * V, 'share_of_step': t_sdf / step_ms}
        B = algorithmic_bytes(V, C, 1 / 3, M2, M3, M4, args.rays, 1 / 3)
        step_roof = {'algorithmic_bytes_per_step': B, 'achieved_gbs': B / (ms / args.steps * 1e-3) / 1e9,
                     'frac_of_hbm': B / (ms / args.steps * 1e-3) / 1e9 / peak, 'M0': M0, 'M2': M2, 'M3': M3, 'M4': M4,
                     'note': 'B = algorithmic bytes of the DENSE formulation (SURVEY 8d); the fused step moves fewer (sparse-aware k0 Adam, '
                             'mask-aware TV), so this is a dense-equivalent rate, not measured traffic'}
        h2d = sum(t.numel() * t.element_size() for t in pool[0])
        line = {'metric': 'rays/sec fwd+bwd+Adam (fine 256^3, 8192-ray batch)', 'value': value, 'unit': 'rays/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
                'clocks': clk.summary(), 'gpu_launches': int(launches),
                'e2e': {'value': total_rays / (ms_e2e * 1e-3), 'unit': 'rays/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                        'ms_per_step': ms_e2e / args.steps},
                'roofline': roof, 'step_roofline': step_roof}
        line.update(extra)
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_reference_arm(args, 1, 1, args.cpu_rays)
        print(json.dumps(line))
    if world > 1:
        if fused is not None and fused._graphs:     # release captured NCCL work before tearing the group down
            fused._graphs.clear()
            torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
