"""Cost of deterministic=True (64-bit fixed-point scatter accumulators) on the headline step: ms/step with and without."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = sys.argv[:1]
import bench as B  # noqa: E402
from voxurf_b200.fused import FusedFineStep  # noqa: E402
from voxurf_b200.trainer import FINE_TRAIN  # noqa: E402

args = B.parse([])
dev = torch.device('cuda', 0)
for det in (False, True):
    m = B.build_model(args, dev)
    fs = FusedFineStep(m, 8192, FINE_TRAIN, B.RENDER_KW, use_graph=True, defer_optimizer=True, deterministic=det)
    pool = [tuple(t.to(dev) for t in b) for b in B.ray_pool(8, 8192, 0)]
    fs.calibrate(*pool[0][:3], global_step=B.START_STEP, headroom=1.35)
    gs = fs.warm_up(pool, B.START_STEP)
    for i in range(B.START_STEP + 20 - gs):
        fs.step(*pool[(gs + i) % 8], gs + i)
    gs = B.START_STEP + 20
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40):
        fs.step(*pool[(gs + i) % 8], gs + i)
    e1.record()
    torch.cuda.synchronize()
    print('deterministic', det, 'ms/step %.4f' % (e0.elapsed_time(e1) / 40), 'M4', fs.counts()[2])
    del fs, m
    torch.cuda.empty_cache()
