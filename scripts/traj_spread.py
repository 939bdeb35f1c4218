"""How far do two runs of the SAME fused step drift apart over the first steps of the benchmarked shape?  (fp32 atomics land in
a different order each run; Adam's first steps from zero moments are sign-like.)  Compares rgb_marched per step between
independent instances: same configuration twice, graph/deferred vs eager, and the deterministic mode twice."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from voxurf_b200.fused import FusedFineStep
from voxurf_b200.trainer import FINE_TRAIN

args = bench.parse([])
N = args.rays
pool = bench.ray_pool(3, N, 0)
dpool = [tuple(t.cuda() for t in b) for b in pool]


def run(**kw):
    torch.manual_seed(0)
    m = bench.build_model(args, torch.device('cuda'))
    fs = FusedFineStep(m, N, FINE_TRAIN, bench.RENDER_KW, **kw)
    fs.calibrate(*dpool[0][:3], global_step=bench.START_STEP, headroom=1.35)
    out = []
    for it in range(0, 8):
        gs = bench.START_STEP + it
        fs.step(*dpool[it % 3], gs)
        fs.apply_lr_decay()
        out.append(fs.rgb_marched.clone())
    fs.sync_params()
    sdf = m.sdf.grid.detach().clone()
    fs.release_graphs()
    return out, sdf


def cmp(a, b, name):
    fr = []
    for x, y in zip(a[0], b[0]):
        d = (x - y).abs()
        fr.append('%.4f/%.1e' % (float((d > 1e-4).float().mean()), float(d.max())))
    ds = (a[1] - b[1]).abs()
    print(name, ' '.join(fr), '| sdf frac>1e-4 %.2e max %.1e' % (float((ds > 1e-4).float().mean()), float(ds.max())), flush=True)


A = run(use_graph=True, defer_optimizer=True)
B = run(use_graph=True, defer_optimizer=True)
C = run(use_graph=False, defer_optimizer=False)
D = run(use_graph=False, defer_optimizer=False, deterministic=True)
E = run(use_graph=True, defer_optimizer=True, deterministic=True)
F = run(use_graph=False, defer_optimizer=False, sparse_adam=False)
print('per step: fraction of rgb values differing by > 1e-4 / max difference')
cmp(A, B, 'graph+defer vs graph+defer  ')
cmp(A, C, 'graph+defer vs eager        ')
cmp(C, D, 'eager vs deterministic eager')
cmp(D, E, 'determ. eager vs determ. graph+defer')
cmp(C, F, 'eager vs eager dense Adam   ')
cmp(D, F, 'determ. vs eager dense Adam ')
