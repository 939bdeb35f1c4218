"""bench.py --workload coarse: BASELINE config 2 -- the 96^3 coarse-stage training step (lib/voxurf_coarse.py:513-619,
run.py:600-659, configs/dtu_e2e/coarse.py) through voxurf_b200.fused_coarse.FusedCoarseStep (one CUDA-graph replay per
step), `--path dropin` = the autograd mirror + Trainer.  Metric: iterations/s on one B200 (ranks > 1 run replicas)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from voxurf_b200 import synthetic as S  # noqa: E402

G, WIDTH, START = 96, 128, 2001
RK = dict(near=0.3, far=6.0, bg=0.0, stepsize=0.5)


def main(args, rank, world, local_rank):
    import bench as B
    C, N = args.k0_channels, args.rays
    config = {'workload': f'voxurf_coarse fwd+bwd+TV+Adam (config 2): {G}^3 SDF + {C}-ch k0, rgbnet 57->128->128->3, per-iteration 5^3 smoothing, '
                          f'ori_tv every iteration, {N}-ray batch, synthetic sphere scene, step {START}+', 'grid': G, 'k0_channels': C,
              'rays_per_gpu': N, 'l2_policy': 'the whole coarse state (grids + moments ~ 140 MB) is of the order of L2; the step is latency / MLP bound',
              'parallelism': f'{world} independent replicas'}
    metric = 'iterations/sec, coarse stage (96^3, 8192-ray batch)'
    from tests.helpers import T
    sc = S.make_coarse_scene(G, C, WIDTH, seed=0, mask_G=48)
    pool = []
    for b in range(4):
        o, d, v = S.make_rays(N, seed=900 + b)
        pool.append(tuple(T(x) for x in (o, d, v, S.make_target(v, seed=b))))
    if args.impl == 'reference':
        if rank != 0:
            return
        from oracle import voxurf_ref as R
        from tests.helpers import oracle_coarse_model
        from voxurf_b200.trainer import COARSE_TRAIN as c
        import torch.nn.functional as F
        torch.set_num_threads(os.cpu_count())
        om = oracle_coarse_model(sc)
        params = [om['sdf'], om['k0']] + [t for W, b in om['rgbnet'] for t in (W, b)]
        lrs = [c['lrate_sdf'], c['lrate_k0']] + [c['lrate_rgbnet']] * 6
        state = [(torch.zeros_like(p), torch.zeros_like(p)) for p in params]
        ts = []
        n_steps = max(1, min(args.steps, 5))
        for it in range(n_steps + 1):
            t0 = time.perf_counter()
            for p in params:
                p.grad = None
            o, d, v, tg = pool[it % 4]
            ret = R.coarse_forward(om, o, d, v, START + it, near=0.3, stepsize=0.5, bg=0.0)
            loss = F.mse_loss(ret['rgb_marched'], tg)
            loss = loss + c['weight_tv_density'] * R.smooth_grad_tv(ret['_full_gradient'], om['nonempty_mask'], c['tv_terms']['smooth_grad_tv'])
            loss = loss + c['weight_tv_density'] * (R.total_variation_coarse(om['sdf'], om['nonempty_mask']) / 2 / om['voxel_size'] * c['tv_terms']['sdf_tv'])
            loss = loss + c['weight_tv_k0'] * R.total_variation_coarse(om['k0'], om['nonempty_mask'].repeat(1, C, 1, 1, 1))
            loss.backward()
            with torch.no_grad():
                for p, lr, (ea, es) in zip(params, lrs, state):
                    if p.grad is not None:
                        R.python_adam_step(p, p.grad, ea, es, it + 1, lr)
            if it:
                ts.append(time.perf_counter() - t0)
        t = float(np.mean(ts))
        cb = {'value': 1.0 / t, 'unit': 'iterations/s', 'cores': os.cpu_count(), 'kind': 'port', 'sample': f'{len(ts)} full {N}-ray coarse steps, {t:.2f} s/step', 's_per_step': t}
        print(json.dumps({'impl': 'reference', 'metric': metric, 'value': 1.0 / t, 'unit': 'iterations/s', 'n_gpus': args.gpus, 'steps': args.steps,
                          'warmup': args.warmup, 'ms_per_step': 1e3 * t, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                          'dtype': 'f32', 'data': 'synthetic', 'config': config, 'cpu_baseline': cb,
                          'e2e': {'value': 1.0 / t, 'unit': 'iterations/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    from tests.helpers import product_coarse_model
    from voxurf_b200 import _lib
    from voxurf_b200.trainer import COARSE_TRAIN, Trainer
    m = product_coarse_model(sc, device=dev)
    dpool = [tuple(t.to(dev) for t in b) for b in pool]
    ppool = [tuple(t.pin_memory() for t in b) for b in pool]
    if args.path == 'fused':
        from voxurf_b200.fused_coarse import FusedCoarseStep
        fs = FusedCoarseStep(m, N, COARSE_TRAIN, RK, use_graph=not args.no_graph)
        fs.calibrate(*dpool[0][:3], global_step=START, headroom=1.5)
        step_fn = lambda b, gs: fs.step(*b, gs)
        decay = fs.apply_lr_decay
    else:
        tr = Trainer(m, COARSE_TRAIN, RK, zero_grad_in_step=False)
        step_fn = lambda b, gs: tr.step(*b, global_step=gs)[0]
        decay = lambda: None
    gs = [START]

    def run(n, e2e=False):
        host = torch.empty(1, pin_memory=True)
        for _ in range(n):
            b = tuple(t.to(dev, non_blocking=True) for t in ppool[gs[0] % 4]) if e2e else dpool[gs[0] % 4]
            loss = step_fn(b, gs[0])
            decay()
            gs[0] += 1
            if e2e:
                host.copy_(loss.reshape(1))
    run(max(args.warmup, 3) + 3)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    r0 = fs.launches_replayed if args.path == 'fused' else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(args.steps); e1.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - l0 + ((fs.launches_replayed - r0) if args.path == 'fused' else 0)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(); run(args.steps, True); f1.record()
    torch.cuda.synchronize()
    ms, ms_e2e = e0.elapsed_time(e1) / args.steps, f0.elapsed_time(f1) / args.steps
    if rank == 0:
        line = {'metric': metric, 'value': 1e3 / ms * world, 'unit': 'iterations/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': config, 'execution': {'path': args.path, 'cuda_graph': bool(args.path == 'fused' and not args.no_graph)},
                'gpu_launches': int(launches), 'rays_per_sec': N * 1e3 / ms * world,
                'e2e': {'value': 1e3 / ms_e2e * world, 'unit': 'iterations/s', 'h2d_bytes_per_step': sum(t.numel() * 4 for t in pool[0]),
                        'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e}}
        if args.path == 'fused':
            M0, M2, M4 = fs.counts()
            line['counts'] = {'M0': M0, 'M2': M2, 'M4': M4}
            fl = (fs.mlp.tc_fwd.flops_per_row() + fs.mlp.tc_bwd.flops_per_row() + fs.mlp.tc_fwd.flops_per_row()) * M4
            tf_peak, tf_src = B.measured_tensor_peak()
            line['mlp_flops_per_step'] = fl
            line['roofline'] = {'bound': 'tensor', 'kernel': 'whole step vs its MLP flops (the step is latency / tensor bound: the grids are 96^3)',
                                'achieved': fl / (ms * 1e-3) / 1e12, 'peak': tf_peak, 'peak_source': tf_src, 'unit': 'TFLOP/s',
                                'frac': fl / (ms * 1e-3) / 1e12 / tf_peak, 'traffic': None}
        print(json.dumps(line))
