#!/usr/bin/env python
"""bench.py -- Voxurf ray-batch volume-rendering hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train|coarse|render|mesh]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline metric (BASELINE.json): rays/sec for fwd + bwd + TV + per-voxel Adam on one 8192-ray batch per GPU per step,
fine stage, 256^3 grids (`--workload train`, the default).  The other workloads are BASELINE.json's remaining configs:
coarse = config 2 (96^3 coarse stage, iterations/s), render = config 4 (forward-only 800x800 views in 8192-ray chunks),
mesh = config 5 (512^3 SDF + gradient field in X-slabs).  Prints ONE JSON line (rank 0).  DESIGN.md section 5 explains
every field.

Arms
  ours       the CUDA path of this repo through its public API (voxurf_b200.fused.FusedFineStep.step, one CUDA-graph
             replay per step on one GPU).
  reference  the reference's own formulation of the same step restated for CPU (oracle/voxurf_ref.py +
             oracle/ref_kernels.c), every timed step the FULL batch on the host cores of this box.
"""
import argparse
import glob
import hashlib
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
METRIC = 'rays/sec fwd+bwd+Adam (fine 256^3, 8192-ray batch)'
PRE_STEPS = 12       # untimed steps before the --warmup steps: eager first occurrences + CUDA-graph captures of every step variant


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=6)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='train', choices=['train', 'coarse', 'render', 'mesh'])
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: --rays per GPU; strong: --rays split over the GPUs (SURVEY 8e asks for both)')
    ap.add_argument('--k0-channels', type=int, default=12, help='12 = BASELINE.json / default_fine_s.py:111; 6 = dtu_e2e/fine.py:73')
    ap.add_argument('--grid', type=int, default=G_FINE)
    ap.add_argument('--rays', type=int, default=N_RAYS)
    ap.add_argument('--smooth', type=int, default=0, help='per-iteration Gaussian smoothing ksize (0 = reference fine config)')
    ap.add_argument('--layout', default='channels_last', choices=['channels_last', 'channel_major'])
    ap.add_argument('--cpu-steps', type=int, default=2, help='timed full-batch CPU steps of the cpu_baseline leg (after 1 warm-up)')
    ap.add_argument('--cpu-rays', type=int, default=0, help='(tests only) shrink the CPU arm\'s batch; 0 = the full batch')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--path', default='fused', choices=['fused', 'dropin'], help='fused = sync-free FusedFineStep; dropin = Voxurf.forward + autograd')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel of the step separately (no CUDA-graph replay)')
    ap.add_argument('--no-defer', action='store_true', help='run the optimizer phase at the end of its own step instead of beside the next step\'s march')
    ap.add_argument('--dense-adam', action='store_true', help='k0 Adam over every voxel (no touched/live bitmaps)')
    ap.add_argument('--dense-exchange', action='store_true', help='multi-GPU: plain dense all-reduces instead of the slab-sharded exchange')
    ap.add_argument('--no-k0-ownership', action='store_true', help='multi-GPU: every rank re-scatters all k0 rows and steps every voxel (no peer-memory stores)')
    ap.add_argument('--sustain', type=float, default=2.0, help='seconds of back-to-back steps for the `sustained` figure (0 = skip)')
    ap.add_argument('--no-parity-check', action='store_true', help='N > 1: skip the untimed replica / gradient parity check')
    ap.add_argument('--phases', action='store_true', help='also print the per-kernel CUDA-event breakdown to stderr')
    ap.add_argument('--views', type=int, default=2, help='render workload: 800x800 views per step-group')
    ap.add_argument('--mesh-res', type=int, default=512)
    return ap.parse_args(argv)


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

    def __init__(self, index, period_ms=20):
        self.rows, self.proc, self.index, self.period = [], None, index, period_ms
        self.marks = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', str(self.period)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(',')]))

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0=None, t1=None):
        sm, mx, reasons = [], [], set()
        for t, r in self.rows:
            if t0 is not None and not (t0 <= t <= t1):
                continue
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
def _peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        return json.load(open(p))
    except Exception:
        return {}


def measured_peak():
    pk = _peaks()
    if 'hbm_gbs' in pk:
        return float(pk['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def measured_tensor_peak():
    pk = _peaks()
    if 'bf16_tflops_sustained' in pk:
        return float(pk['bf16_tflops_sustained']), 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)'
    return 1590.0, 'fallback (B200_PROFILING.md, 1.59 PFLOP/s bf16)'


def algorithmic_bytes(V, C, tv, M2, M3, M4, N, fd):
    """B = 4V[9(1+C) + 3 tv + 4 fd] + (16 M2 + 8 M3 + 20 M4) + 60 N"""
    return 4 * V * (9 * (1 + C) + 3 * tv + 4 * fd) + (16 * M2 + 8 * M3 + 20 * M4) + 60 * N


def source_sha():
    """Hash of the CUDA sources the loaded library was built from: profiles/traffic.json is keyed by it, so a dram-traffic
    figure measured on another build is never printed."""
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, 'voxurf_b200', 'csrc', '*.cu*'))):
        h.update(open(f, 'rb').read())
    return h.hexdigest()[:16]


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the ncu --set full capture of THIS build
    (profiles/traffic.json, written by scripts/ncu_traffic.py), else None."""
    try:
        t = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
    except Exception:
        return None
    if t.get('src_sha') != source_sha():
        return None
    v = t.get('kernels', {}).get(kernel)
    return float(v) if v is not None else None


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def oracle_bench_model(G, C, smooth, k0=None, mlps=None, sdf=None):
    """The bench scene as the oracle's model dict (TEST / BASELINE infrastructure: only the reference arm, cpu_baseline and
    tests call this).  k0 / mlps: take these tensors instead of the seeded defaults (parity tests pass the product's)."""
    from oracle import voxurf_ref as R
    xyz_min, xyz_max = torch.tensor([-1., -1., -1.]), torch.tensor([1., 1., 1.])
    cfg = S.FINE_CFG
    rs = np.random.RandomState(0)
    dim0, k_dim0 = S.fine_dims(C)
    lay = lambda ls: [(torch.as_tensor(W).clone().requires_grad_(True), torch.as_tensor(b).clone().requires_grad_(True)) for W, b in ls]
    mc = dict(density=R.mask_cache_density(torch.from_numpy(S.mask_density(100))), xyz_min=xyz_min, xyz_max=xyz_max,
              act_shift=float(np.log(1 / (1 - 1e-6) - 1)), voxel_size_ratio=1.0, thres=1e-3)
    voxel_size = ((xyz_max - xyz_min).prod() / (G ** 3)).pow(1 / 3)
    nonempty = R.nonempty_mask(mc, xyz_min, xyz_max, (G, G, G))
    if sdf is None:
        sdf = torch.from_numpy(S.sphere_sdf(G))
        sdf[~nonempty] = 1
    sdf = sdf.clone()
    if k0 is None:
        k0 = 0.1 * torch.randn(1, C, G, G, G, generator=torch.Generator().manual_seed(1234))
    if mlps is None:
        mlps = (S.mlp_init(rs, dim0, WIDTH, 4), S.mlp_init(rs, k_dim0, WIDTH, 4))
    m = dict(xyz_min=xyz_min, xyz_max=xyz_max, voxel_size=voxel_size, mask_cache=mc, nonempty_mask=nonempty,
             sdf=sdf.requires_grad_(True), k0=k0.clone().requires_grad_(True), rgbnet=lay(mlps[0]), k_rgbnet=lay(mlps[1]),
             posfreq=torch.FloatTensor([2 ** i for i in range(5)]), viewfreq=torch.FloatTensor([1.]),
             k_posfreq=torch.FloatTensor([2 ** i for i in range(5)]), k_viewfreq=torch.FloatTensor([1.]),
             grad_feat=cfg['grad_feat'], use_grad_norm=True, center_sdf=True, k_center_sdf=False, k_res=True,
             fast_color_thres=1e-4, s_ratio=50, s_start=0.05, step_start=0, s_val=0.05,
             smooth_kernel=R.gaussian_kernel3d(smooth, 0.8) if smooth > 0 else None)
    params = [m['sdf'], m['k0']] + [t for W, b in m['rgbnet'] for t in (W, b)] + [t for W, b in m['k_rgbnet'] for t in (W, b)]
    lrs = [5e-3, 1e-1] + [3e-3] * 8 + [1e-3] * 8
    state = [(torch.zeros_like(p), torch.zeros_like(p)) for p in params]
    return m, params, lrs, state


def oracle_bench_step(m, params, lrs, state, batch, gs, adam_step, n_batch, G, lr_scale=1.0, grads_out=None, grads_out_tv=None):
    """One full training iteration of the reference's formulation on CPU: forward + losses (run.py:604-636), backward,
    on TV iterations the smooth-gradient TV and the TV add-grad (run.py:612-655), the dense python Adam (lib/utils.py).
    -> (loss incl. regulariser, ret dict)"""
    from oracle import kernels as K
    from oracle import voxurf_ref as R
    o, d, v, target = batch
    for p in params:
        p.grad = None
    ret = R.fine_forward(m, o, d, v, gs, near=RENDER_KW['near'], stepsize=0.5, bg=RENDER_KW['bg'])
    loss = R.fine_loss(ret, target)
    tv_iter = gs % 3 == 0
    if tv_iter:
        loss = loss + 0.01 * R.smooth_grad_tv(ret['_full_gradient'], m['nonempty_mask'], 0.05)
    loss.backward()
    if grads_out is not None:    # parity tests: gradients of the data + smooth-grad-TV terms, before the TV add-grad
        grads_out.extend(None if p.grad is None else p.grad.clone() for p in params)
    if tv_iter:
        w = 0.01 * 0.1 / n_batch * G / 128
        K.total_variation_add_grad(m['sdf'].detach(), m['sdf'].grad, w, w, w, True)
    if grads_out_tv is not None:  # parity tests: the gradients the optimizer consumes (after the TV add-grad)
        grads_out_tv.extend(None if p.grad is None else p.grad.clone() for p in params)
    with torch.no_grad():
        for p, lr, (ea, es) in zip(params, lrs, state):
            if p.grad is not None:
                R.python_adam_step(p, p.grad, ea, es, adam_step, lr * lr_scale)
    return loss.detach(), ret


def cpu_reference_arm(args, steps, warmup, budget_s=150.0):
    """The oracle's restatement of the step, timed on the host cores: every step is the full batch (no extrapolation).
    A full-batch step takes 2.5 - 6 s on the boxes' host CPUs, so the run is bounded by `budget_s` of wall clock: at most
    `warmup` warm-up steps inside the first quarter of the budget (at least one), then up to `steps` timed steps while the
    budget lasts (at least two); the `sample` text says how many were timed."""
    torch.set_num_threads(os.cpu_count())
    G, C = args.grid, args.k0_channels
    n_rays = args.cpu_rays or args.rays
    m, params, lrs, state = oracle_bench_model(G, C, args.smooth)
    pool = ray_pool(4, n_rays, 0)
    ts, it, n_warm = [], 0, 0
    t_start = time.perf_counter()
    while len(ts) < steps:
        timed = n_warm >= warmup or (n_warm >= 1 and time.perf_counter() - t_start > budget_s / 4)
        if timed and len(ts) >= min(2, steps) and time.perf_counter() - t_start > budget_s:
            break
        t0 = time.perf_counter()
        oracle_bench_step(m, params, lrs, state, pool[it % len(pool)], START_STEP + it, it + 1, n_rays, G)
        it += 1
        if timed:
            ts.append(time.perf_counter() - t0)
        else:
            n_warm += 1
    t = float(np.mean(ts))
    return {'value': n_rays / t, 'unit': 'rays/s', 'cores': os.cpu_count(), 'kind': 'port',
            'sample': f'{len(ts)} full {n_rays}-ray steps (fwd + bwd + TV every 3rd + dense Adam over {G}^3 x (1+{C}) + MLPs) after {n_warm} '
                      f'warm-up, {t:.2f} s/step, torch {torch.get_num_threads()} threads; oracle/voxurf_ref.py + oracle/ref_kernels.c'
                      + (f' (requested {steps} timed / {warmup} warm-up steps: bounded to ~{budget_s:.0f} s of CPU work)' if len(ts) < steps or n_warm < warmup else ''),
            's_per_step': t, 'steps_timed': len(ts)}


# ------------------------------------------------------------------------------------------------ main
def common_config(args, world):
    C, G = args.k0_channels, args.grid
    rays = args.rays if args.scaling == 'weak' else args.rays // world
    workload = (f'voxurf_fine fwd+bwd+TV+Adam, {G}^3 SDF + {C}-ch k0, rgbnet 79->192^3->3 + k_rgbnet, {rays}-ray batch/GPU, '
                f'synthetic sphere scene, step {START_STEP}+, TV every 3rd iter, smooth_ksize={args.smooth}')
    return {'workload': workload, 'grid': G, 'k0_channels': C, 'rays_per_gpu': rays, 'k0_layout': args.layout,
            'l2_policy': 'inputs larger than L2 (grids + moments 0.27 GB x (1+C) >> 126 MB)', 'parallelism': f'dp{world} over rays'}


def main():
    args = parse()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.workload != 'train':
        import bench_workloads as BW   # coarse stage / render-only / mesh-field workloads (BASELINE configs 2, 4, 5)
        return BW.main(args, rank, world, local_rank)
    C, G = args.k0_channels, args.grid
    config = common_config(args, world)
    rays = config['rays_per_gpu']

    if args.impl == 'reference':
        if rank != 0:
            return
        cb = cpu_reference_arm(args, max(1, args.steps), max(0, args.warmup))
        line = {'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': 'rays/s',
                'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * cb['s_per_step'],
                'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
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
    pool = ray_pool(8, rays, rank)
    dev_pool = [tuple(t.to(device) for t in b) for b in pool]
    pin_pool = [tuple(t.pin_memory() for t in b) for b in pool]
    fused = None
    if args.path == 'fused':
        from voxurf_b200.fused import FusedFineStep
        fused = FusedFineStep(model, rays, FINE_TRAIN, RENDER_KW, world=world, rank=rank, sparse_adam=not args.dense_adam,
                              use_graph=not args.no_graph, dense_exchange=args.dense_exchange, defer_optimizer=not args.no_defer,
                              k0_ownership=not args.no_k0_ownership)
        fused.calibrate(*dev_pool[0][:3], global_step=START_STEP, headroom=1.35)
        step_fn = lambda b, gs: fused.step(*b, gs)
        decay = fused.apply_lr_decay
    else:
        sync = parallel.GradSync(model, world) if world > 1 else None
        trainer = Trainer(model, FINE_TRAIN, RENDER_KW, zero_grad_in_step=False, grad_sync=sync)
        trainer.global_batch = rays * world
        step_fn = lambda b, gs: trainer.step(*b, global_step=gs)[0]
        f_decay = 0.1 ** (1 / (FINE_TRAIN['lrate_decay'] * 1000))

        def decay():   # run.py:679-683
            for g in trainer.optimizer.param_groups:
                g['lr'] *= f_decay

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gs_next = [START_STEP]

    def run(n, e2e=False):
        loss_host = torch.empty(1, pin_memory=True)
        for _ in range(n):
            gs = gs_next[0]
            gs_next[0] += 1
            b = gs % len(pool)
            batch = tuple(t.to(device, non_blocking=True) for t in pin_pool[b]) if e2e else dev_pool[b]
            loss = step_fn(batch, gs)
            decay()
            if e2e:
                loss_host.copy_(loss.reshape(1), non_blocking=False)

    # ---- warm-up: every execution variant (TV / non-TV) run eagerly once and captured BEFORE anything is timed
    # (the clock sampler starts here: nvidia-smi needs a few hundred ms before its first sample)
    clk = ClockSampler(local_rank, period_ms=10)
    clk.__enter__()
    if fused is not None:
        gs_next[0] = fused.warm_up(dev_pool, START_STEP)
        run(max(0, START_STEP + PRE_STEPS - gs_next[0]))     # the timed window starts at the same global step in every configuration
    run(max(args.warmup, 3))
    barrier()
    graph = fused is not None and fused.use_graph
    execution = {'path': args.path, 'cuda_graph': bool(graph), 'graphs_captured': len(fused._graphs) if graph else 0,
                 'optimizer': ('deferred: the Adam / regulariser phase of step k runs beside the march of step k + 1 (one optimizer phase per '
                               'timed step all the same)') if (fused is not None and fused.defer_optimizer) else 'end of step'}

    if fused is not None and world > 1:
        execution['exchange'] = (('sdf: owners pull the non-zero 128-voxel gradient blocks of every rank over NVLink peer memory (vx_pull_reduce), slab '
                                  'Adam stores the updated blocks into the peers\' replicas (vx_adam_step_blocklive_peers); ' if fused.sdf_peer else
                                  'sdf: in-place reduce-scatter over X-slabs + slab Adam + in-place all-gather of the parameters; ') +
                                 'MLPs: one all-reduce; k0: rows all-gathered, ' +
                                 ('each rank scatters / steps its own X-slab and stores the updated voxels into the peers\' replicas over '
                                  'NVLink peer memory (vx_adam_step_worklist_peers)' if fused.k0_owned else
                                  'every rank re-scatters all rows and steps every voxel' + (f' [{fused.k0_peer_note}]' if fused.k0_peer_note else ''))
                                 ) if fused.sharded else 'dense all-reduces (sdf, MLPs) + k0 row exchange'
    try:
        # ---- device-resident timing
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        r0 = fused.launches_replayed if graph else 0
        tw0 = time.perf_counter()
        ev0.record()
        run(args.steps)
        ev1.record()
        barrier()
        tw1 = time.perf_counter()
        launches = _lib.launch_count() - l0 + ((fused.launches_replayed - r0) if graph else 0)
        ms = ev0.elapsed_time(ev1)
        # ---- end-to-end timing (pinned host -> device inputs, device -> host loss, every step)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(args.steps, e2e=True)
        e1.record()
        barrier()
        tw2 = time.perf_counter()
        ms_e2e = e0.elapsed_time(e1)
        # ---- sustained: >= --sustain seconds of back-to-back steps (clocks settle; MEASURED_PEAKS shows this part drops
        # to ~1.4 GHz under sustained tensor load)
        sustained = None
        if args.sustain > 0:
            t_ = torch.tensor([ms], device=device, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t_, op=dist.ReduceOp.MAX)     # every rank must run the same number of steps
            n_sus = max(args.steps, int(args.sustain / max(float(t_[0]) / args.steps * 1e-3, 1e-6)))
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            ts0 = time.perf_counter()
            s0.record()
            run(n_sus)
            s1.record()
            barrier()
            ts1 = time.perf_counter()
            sustained = {'steps': n_sus, 'ms_per_step': s0.elapsed_time(s1) / n_sus}
    finally:
        clk.__exit__()
    t = torch.tensor([ms, ms_e2e, sustained['ms_per_step'] if sustained else 0.0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if sustained:
        sustained['ms_per_step'] = float(t[2])
        sustained['value'] = rays * world / (sustained['ms_per_step'] * 1e-3)
        sustained['clocks'] = clk.summary(ts0, ts1)

    # ---- per-kernel event timing: the same step launched kernel by kernel (no graph) right after, outside the headline
    kt = None
    if fused is not None:
        fused.force_eager = True
        run(3)
        barrier()
        _lib.TIMING, fused.timings = [], []
        n_k = 6
        run(n_k)
        barrier()
        rows, _lib.TIMING = _lib.TIMING, None
        t_adam = {}
        for name, ev in fused.timings:
            t_adam.setdefault(name, []).append(ev[0].elapsed_time(ev[1]))
        fused.timings = None
        fused.force_eager = False
        kt = {}
        for name, a, b in rows:
            kt.setdefault(name, []).append(a.elapsed_time(b))
        kt = {k: (sum(v) / n_k, len(v) / n_k) for k, v in kt.items()}   # (ms per step, launches per step)
        if args.phases and rank == 0:
            for k, (tms, n) in sorted(kt.items(), key=lambda x: -x[1][0]):
                print(f'{tms * 1e3:9.1f} us/step  {n:5.2f} launches/step  {k}', file=sys.stderr)

    parity = None
    if world > 1 and fused is not None and not args.no_parity_check:
        parity = fused.parity_check(dev_pool[0], gs_next[0])

    probe = None
    if fused is not None and fused.k0_touched is not None:
        # one extra (untimed) step on every rank to count the voxels the sparse-aware k0 Adam pass touches
        fused.bitmap_probe = []
        fused.force_eager = True
        run(1)
        torch.cuda.synchronize()
        fused.force_eager = False
        probe, fused.bitmap_probe = fused.bitmap_probe, None
    if fused is not None:
        fused.sync_params()
        fused.poll_overflow(force=True)
    if rank == 0:
        total_rays = rays * world * args.steps
        value = total_rays / (ms * 1e-3)
        clocks = clk.summary(tw0, tw2)          # the device-resident and end-to-end timed regions
        if clocks['samples'] < 3 and sustained:
            # 2 x 30 steps are shorter than three nvidia-smi samples: take the samples of the whole measurement (timed, e2e and
            # the sustained repeat of the same steps, back to back)
            clocks = clk.summary(tw0, ts1)
            clocks['window'] = 'timed + e2e + sustained regions'
        V = G ** 3
        peak, peak_src = measured_peak()
        tf_peak, tf_src = measured_tensor_peak()
        step_ms = ms / args.steps
        line = {'metric': METRIC, 'value': value, 'unit': 'rays/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': step_ms, 'higher_is_better': True, 'scaling': args.scaling,
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config, 'execution': execution,
                'clocks': clocks, 'gpu_launches': int(launches),
                'e2e': {'value': total_rays / (ms_e2e * 1e-3), 'unit': 'rays/s',
                        'h2d_bytes_per_step': sum(t_.numel() * t_.element_size() for t_ in pool[0]), 'd2h_bytes_per_step': 4,
                        'ms_per_step': ms_e2e / args.steps}}
        if sustained:
            line['sustained'] = sustained
        if parity is not None:
            line['parity_check'] = parity
        if fused is not None:
            M0, M2, M4 = fused.counts()
            M3 = int(fused.keep[:M2].sum())
            # ---- dominant kernel by event time, and the two HBM-bound Adam launches
            flops_row = sum(2 * sum(k * n for k, n in zip(c.K, c.N)) for mlp in (fused.mlp1, fused.mlp2) for c in (mlp.tc_fwd, mlp.tc_bwd))
            mlp_names = [k for k in kt if k.startswith('vx_mlp_chain')]
            t_chain = sum(kt[k][0] for k in mlp_names)
            n_chain = sum(kt[k][1] for k in mlp_names)
            fl_step = flops_row * M4                                  # fwd + dX chains of both networks, fp32-equivalent flops
            roof = {'bound': 'tensor',
                    'kernel': 'k_mlp_chain (fused 4-layer MLP chains on tcgen05, TF32x3; %d rows, %.0f launches/step)' % (M4, n_chain),
                    'achieved': fl_step / (t_chain * 1e-3) / 1e12, 'peak': tf_peak, 'peak_source': tf_src, 'unit': 'TFLOP/s',
                    'traffic': measured_traffic('k_mlp_chain'), 'ms_per_launch': t_chain / max(n_chain, 1),
                    'algorithmic_flops_per_launch': fl_step / max(n_chain, 1), 'issued_tf32_flops_per_launch': 3 * fl_step / max(n_chain, 1),
                    'formulation_ceiling_frac': 1.0 / 6.0, 'share_of_step': t_chain / sum(v[0] for v in kt.values())}
            roof['frac'] = roof['achieved'] / tf_peak
            line['roofline'] = roof
            # weight gradients (same tensor bound, reported beside the chains)
            if 'vx_mlp_dw_batch' in kt:
                fl_dw = sum(2 * sum(k * n for k, n in zip(c.K, c.N)) for mlp in (fused.mlp1, fused.mlp2) for c in (mlp.tc_fwd,)) * M4
                t_dw = kt['vx_mlp_dw_batch'][0]
                line['roofline_mlp_dw'] = {'bound': 'tensor', 'kernel': 'k_mlp_dw (8 weight/bias-gradient GEMMs, one launch)',
                                           'achieved': fl_dw / (t_dw * 1e-3) / 1e12, 'peak': tf_peak, 'unit': 'TFLOP/s',
                                           'frac': fl_dw / (t_dw * 1e-3) / 1e12 / tf_peak, 'ms_per_launch': t_dw,
                                           'traffic': measured_traffic('k_mlp_dw')}
            sparse = None
            adam_bytes = 32 * V * C
            if probe:
                tb, lb = (np.unpackbits(x.cpu().numpy().view(np.uint8)) for x in probe[0])
                n_t, n_l = int(tb.sum()), int((tb | lb).sum())
                adam_bytes = 4 * C * (8 * n_t + 6 * (n_l - n_t)) + 2 * 4 * fused.k0_touched.numel()
                sparse = {'voxels': V, 'touched': n_t, 'live': n_l, 'dense_equivalent_bytes': 32 * V * C}
            line['kernel_ms_per_step'] = {k: round(v[0], 5) for k, v in sorted(kt.items(), key=lambda x: -x[1][0])}
            if t_adam.get('k0'):
                tk = float(np.mean(t_adam['k0']))
                line['roofline_k0_adam'] = {'bound': 'hbm', 'kernel': 'k_bitmap_voxel_list + k_adam_voxel_list (k0 grid, voxels in touched | live only)', 'achieved': adam_bytes / (tk * 1e-3) / 1e9,
                                            'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s', 'frac': adam_bytes / (tk * 1e-3) / 1e9 / peak,
                                            'ms_per_launch': tk, 'algorithmic_bytes_per_launch': adam_bytes, 'sparse': sparse,
                                            'traffic': measured_traffic('k_adam_voxel_list')}
            if t_adam.get('sdf'):
                tsd = float(np.mean(t_adam['sdf']))
                n_el = V // (world if fused.sharded else 1)
                bytes_sdf, kname, live_frac = 32 * n_el, 'k_adam (sdf grid, %d^3 fp32, dense, 32 B/element' % G, None
                if fused.sdf_live is not None:
                    # block-live pass: 4 B/element of gradient everywhere, + 28 B/element in the blocks that are (or become) live
                    lv = fused.sdf_live[fused.slab[0] // 128:fused.slab[1] // 128] if fused.sharded else fused.sdf_live
                    live_frac = float(lv.float().mean())
                    bytes_sdf = int(n_el * (4 + 28 * live_frac))
                    kname = 'k_adam_blocklive (sdf grid, %d^3 fp32, 128-voxel blocks: 4 B/element + 28 B/element where live' % G
                line['roofline_sdf_adam'] = {'bound': 'hbm', 'kernel': kname + (', X-slab of this rank)' if fused.sharded else ')'),
                                             'achieved': bytes_sdf / (tsd * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                                             'frac': bytes_sdf / (tsd * 1e-3) / 1e9 / peak, 'ms_per_launch': tsd, 'live_block_fraction': live_frac,
                                             'algorithmic_bytes_per_launch': bytes_sdf,
                                             'traffic': measured_traffic('k_adam_blocklive' if fused.sdf_live is not None else 'k_adam')}
            # ---- step-level bounds: the dense-equivalent HBM figure of SURVEY 8d, and the honest combined bound
            B_dense = algorithmic_bytes(V, C, 1 / 3, M2, M3, M4, rays, 1 / 3)
            n_act = int(fused.tv_active.sum()) if getattr(fused, 'tv_active', None) is not None else V
            # bytes this formulation must move: dense sdf Adam (32 B/voxel), sparse k0 Adam, mask-aware TV passes on TV
            # iterations (FD gradient 4 read + 12 write, smooth-grad 12 + 12, adjoint 12 + 8 per active voxel), sample lists,
            # MLP row images (X + 3 hidden activations + their gradients, written once and read once by the dW pass)
            act_bytes = M4 * 4 * (fused.ld1 + fused.ld2 + 6 * WIDTH) * 2 * 2
            B_actual = 32 * V + adam_bytes + (1 / 3) * 60 * n_act + (16 * M2 + 8 * M3 + 20 * M4) + 60 * rays + act_bytes
            F_step = 1.5 * fl_step                                                   # + weight gradients
            t_min = B_actual / (peak * 1e9) + F_step / (tf_peak * 1e12)
            t_min_form = B_actual / (peak * 1e9) + 6 * F_step / (tf_peak * 1e12)
            line['step_roofline'] = {
                'algorithmic_bytes_per_step_dense': B_dense, 'frac_of_hbm_dense_equivalent': B_dense / (step_ms * 1e-3) / 1e9 / peak,
                'bytes_per_step_this_formulation': B_actual, 'mlp_flops_per_step': F_step,
                't_min_ms': t_min * 1e3, 'frac_of_t_min': t_min * 1e3 / step_ms,
                't_min_tf32x3_ms': t_min_form * 1e3, 'frac_of_t_min_tf32x3': t_min_form * 1e3 / step_ms,
                'M0': M0, 'M2': M2, 'M3': M3, 'M4': M4,
                'note': 't_min = B/BW_hbm + F/peak_tc (SURVEY 8d) with B = bytes this formulation must move (sparse-aware k0 Adam, '
                        'mask-aware TV) and F = MLP fwd + dX + dW flops; the tf32x3 variant charges the 3-pass TF32 split (peak/6). '
                        'frac_of_hbm_dense_equivalent divides the DENSE formulation\'s bytes by the step time: a rate, not traffic'}
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_reference_arm(args, max(1, args.cpu_steps), 1)
        print(json.dumps(line))
    if world > 1:
        if fused is not None:
            fused.shutdown()     # captured NCCL work and the peers' memory mappings must go before the process group does
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
