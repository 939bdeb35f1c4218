"""CPU (gloo, world_size 2): the host-side logic of the ray-sharded data-parallel path -- gradient averaging across
ranks equals the gradient of the concatenated batch, bucketing of small tensors, slab / chunk sharding."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from voxurf_b200.parallel import GradSync, shard_range


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 79, 512):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


class Toy(nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.grid = nn.Parameter(torch.randn(1, 3, 9, 8, 7))        # "large" tensor path (threshold lowered below)
        self.net = nn.Sequential(nn.Linear(5, 4), nn.ReLU(), nn.Linear(4, 3))

    def forward(self, x):
        return (self.net(x).sum(-1) * self.grid.mean() + (self.grid[0, :, 0, 0, 0] * x[:, :3]).sum(-1)).pow(2).mean()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(123)
    x_all = torch.randn(8 * world, 5)
    m = Toy()
    lo, hi = shard_range(x_all.shape[0], rank, world)
    m(x_all[lo:hi]).backward()
    GradSync(m, world, small_numel=100)(m)
    grads = [p.grad.clone() for p in m.parameters()]
    ref = Toy()
    ref(x_all).backward()                                           # batch-mean loss on the concatenated batch
    ok = all(torch.allclose(g, p.grad, rtol=1e-5, atol=1e-7) for g, p in zip(grads, ref.parameters()))
    out[rank] = ok
    dist.destroy_process_group()


def test_gradsync_equals_concatenated_batch_gradient():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)
