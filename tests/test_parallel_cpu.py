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


# ------------------------------------------------------------------------------------------------ the sharded exchange protocol
def _protocol_worker(rank, world, port, out):
    """The slab-ownership exchange of FusedFineStep, restated with gloo collectives standing in for the
    NVLink peer loads / stores of the CUDA kernels: block flags (vx_block_nonzero) -> the slab owner sums, in rank order, the
    flagged 128-element blocks of every rank's gradient and scales by 1 / world (vx_pull_reduce) -> Adam on the owned slab only
    (oracle's lib/utils.py Adam) -> the owner's updated slab lands in every replica (vx_adam_step_blocklive_peers)."""
    from oracle import voxurf_ref as R
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    n_blocks, B = 16 * world, 128
    N = n_blocks * B
    torch.manual_seed(7)
    p0 = torch.randn(N)
    p = p0.clone()
    m_own, v_own = torch.zeros(N), torch.zeros(N)                  # only the owned slab of the moments is ever touched
    p_ref, m_ref, v_ref = p0.clone(), torch.zeros(N), torch.zeros(N)   # dense reference: all-reduce(AVG) + replicated Adam
    lo, hi = rank * N // world, (rank + 1) * N // world
    for step in range(1, 4):
        g = torch.zeros(n_blocks, B)
        gen = torch.Generator().manual_seed(100 * step + rank)
        for b in torch.randperm(n_blocks, generator=gen)[:n_blocks // 3]:
            g[b, torch.randint(0, B, (9,), generator=gen)] = torch.randn(9, generator=gen)
        g = g.reshape(-1)
        # --- protocol
        mask = (g.view(n_blocks, B) != 0).any(1)
        all_g = [torch.empty(N) for _ in range(world)]
        all_m = [torch.empty(n_blocks, dtype=torch.bool) for _ in range(world)]
        dist.all_gather(all_g, g)                                    # (the kernels read only the flagged blocks of the owned slab)
        dist.all_gather(all_m, mask)
        acc = torch.zeros(hi - lo)
        for q in range(world):                                       # rank order: a fixed summation order
            acc = acc + all_g[q][lo:hi] * all_m[q].repeat_interleave(B)[lo:hi]
        g_slab = acc * torch.tensor(1.0 / world, dtype=torch.float32)
        R.python_adam_step(p[lo:hi], g_slab, m_own[lo:hi], v_own[lo:hi], step, 5e-3)
        slabs = [torch.empty(hi - lo) for _ in range(world)]
        dist.all_gather(slabs, p[lo:hi].clone())                     # (peer stores of the updated blocks)
        p = torch.cat(slabs)
        # --- dense reference
        g_avg = g.clone()
        dist.all_reduce(g_avg)
        g_avg /= world
        R.python_adam_step(p_ref, g_avg, m_ref, v_ref, step, 5e-3)
    digest = [torch.empty(N) for _ in range(world)]
    dist.all_gather(digest, p)
    out[rank] = dict(identical=all(torch.equal(digest[0], d) for d in digest),
                     close=bool(torch.allclose(p, p_ref, rtol=1e-6, atol=1e-7)),
                     moved=float((p - p0).abs().max()),
                     own_moments_only=bool((m_own[:lo] == 0).all() and (m_own[hi:] == 0).all() and (m_own[lo:hi] != 0).any()))
    dist.destroy_process_group()


def test_slab_ownership_exchange_protocol_equals_allreduce_plus_replicated_adam():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_protocol_worker, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        assert out[r]['identical'], 'replicas must be bit-identical (the owner computes, everyone stores its values)'
        assert out[r]['close'], 'same parameters as all-reduce(AVG) + Adam on every rank, up to the summation order'
        assert out[r]['moved'] > 1e-3 and out[r]['own_moments_only']
