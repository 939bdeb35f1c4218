"""GPU (needs >= 2 devices; skipped otherwise): ray-sharded data parallelism of the fused step over NCCL.
N-GPU gradients after the exchange == 1-GPU gradients on the concatenated batch, and so are the parameters after
the TV + Adam step."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
RK = dict(near=0.3, far=6.0, bg=0.0, stepsize=0.5)


def _build(dev):
    from voxurf_b200 import synthetic as S
    from tests.helpers import product_fine_model
    sc = S.make_fine_scene(40, 6, 64, seed=9)
    return product_fine_model(sc, device=dev, k0_channels_last=True)


def _batches(world, n_rays):
    from voxurf_b200 import synthetic as S
    from tests.helpers import T
    out = []
    for r in range(world):
        o, d, v = (T(x) for x in S.make_rays(n_rays, seed=50 + r))
        out.append((o, d, v, T(S.make_target(v.numpy(), seed=r))))
    return out


def _worker(rank, world, port, n_rays, step, sparse, ret, dense_exchange=False, graph=False, defer=False):
    import torch.distributed as dist
    from voxurf_b200.fused import FusedFineStep
    from voxurf_b200.parallel import GradSync
    from voxurf_b200.trainer import FINE_TRAIN
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    m = _build(dev)
    fs = FusedFineStep(m, n_rays, FINE_TRAIN, RK, row_capacity=8192, world=world, rank=rank, sparse_k0_exchange=sparse,
                       dense_exchange=dense_exchange, use_graph=graph, defer_optimizer=defer)
    assert fs.sharded == (not dense_exchange)
    b = [t.to(dev) for t in _batches(world, n_rays)[rank]]
    if graph:
        # whole steps (CUDA-graph capture and replay of the step including its NCCL collectives): parameters only
        for it in range(7):
            fs.step(*b, step + it)
        fs.sync_params()
        fs.poll_overflow(force=True)
        assert len(fs._graphs) == (3 if defer else 2), list(fs._graphs)
        ret[rank] = ({}, {'sdf': m.sdf.grid.detach().cpu(), 'k0': m.k0.grid.detach().contiguous().cpu(), 'mlp1': fs.mlp1.flat.detach().cpu()})
        assert fs.k0_owned, fs.k0_peer_note
        fs.shutdown()
        dist.barrier()
        dist.destroy_process_group()
        return
    fs.forward_backward(*b, step)
    fs.grad_sync()
    assert fs.k0_owned == (fs.sharded and sparse) and fs.sdf_peer == fs.k0_owned, fs.k0_peer_note    # NVLink peer memory is there on the B200 boxes
    fs.gather_k0_grad()      # (k0 ownership: each rank scattered the corners inside its own X-slab only)
    if fs.sharded:     # the reduce-scatter left every rank with its X-slab of the averaged sdf gradient
        flat = m.sdf.grid.grad.view(-1)
        dist.all_gather_into_tensor(flat, flat[fs.slab[0]:fs.slab[1]].clone())
    grads = {'sdf': m.sdf.grid.grad.clone().cpu(), 'k0': m.k0.grid.grad.contiguous().clone().cpu(),
             'mlp1': fs.mlp1.flat.grad.clone().cpu(), 'mlp2': fs.mlp2.flat.grad.clone().cpu()}
    fs.regularise(step)
    fs.optimizer_step()
    fs.sync_params()
    params = {'sdf': m.sdf.grid.detach().cpu(), 'k0': m.k0.grid.detach().contiguous().cpu(), 'mlp1': fs.mlp1.flat.detach().cpu()}
    fs.counts()
    ret[rank] = (grads, params)
    fs.shutdown()
    dist.destroy_process_group()


@pytest.mark.parametrize('sparse,dense_exchange', [(True, False), (False, False), (True, True)])
def test_two_gpu_gradients_equal_single_gpu_on_concatenated_batch(sparse, dense_exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    from voxurf_b200.fused import FusedFineStep
    from voxurf_b200.trainer import FINE_TRAIN
    world, n_rays, step = 2, 512, 15003
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, n_rays, step, sparse, ret, dense_exchange), nprocs=world, join=True)
    # single GPU, concatenated batch
    dev = torch.device('cuda', 0)
    m = _build(dev)
    fs = FusedFineStep(m, n_rays * world, FINE_TRAIN, RK, row_capacity=16384)
    bs = _batches(world, n_rays)
    cat = [torch.cat([b[i] for b in bs]).to(dev) for i in range(4)]
    fs.forward_backward(*cat, step)
    ref = {'sdf': m.sdf.grid.grad.clone().cpu(), 'k0': m.k0.grid.grad.contiguous().clone().cpu(),
           'mlp1': fs.mlp1.flat.grad.clone().cpu(), 'mlp2': fs.mlp2.flat.grad.clone().cpu()}
    fs.regularise(step)
    fs.optimizer_step()
    refp = {'sdf': m.sdf.grid.detach().cpu(), 'k0': m.k0.grid.detach().contiguous().cpu(), 'mlp1': fs.mlp1.flat.detach().cpu()}
    for r in range(world):
        grads, params = ret[r]
        for k in ref:
            scale = float(ref[k].abs().max())
            np.testing.assert_allclose(grads[k].numpy(), ref[k].numpy(), rtol=1e-4, atol=1e-4 * scale, err_msg=f'rank {r} grad {k}')
        for k, lr in (('sdf', 5e-3), ('k0', 1e-1), ('mlp1', 3e-3)):
            np.testing.assert_allclose(params[k].numpy(), refp[k].numpy(), rtol=1e-4, atol=2e-2 * lr, err_msg=f'rank {r} param {k}')
    # both ranks hold identical replicas after the step (bit-identical with the dense all-reduce; the row exchange
    # re-scatters with fp32 atomics whose order differs per rank, so k0 agrees to rounding there)
    # (with k0 ownership -- sharded + row exchange -- the owner's values are stored into every replica: bit-identical too)
    for k in ('sdf', 'mlp1') + (('k0',) if (not sparse or not dense_exchange) else ()):
        assert torch.equal(ret[0][1][k], ret[1][1][k]), k
    np.testing.assert_allclose(ret[0][1]['k0'].numpy(), ret[1][1]['k0'].numpy(), rtol=1e-4, atol=2e-3)


@pytest.mark.parametrize('defer', [False, True])
def test_two_gpu_graph_replayed_steps_follow_single_gpu(defer):
    """Seven whole steps (TV iterations included) as CUDA-graph replays with the slab-sharded exchange captured inside,
    against a single GPU stepping on the concatenated batch.  defer: as bench.py runs it -- the optimizer phase and the
    gradient exchange of step k execute inside the launch of step k + 1, beside its march."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    from voxurf_b200.fused import FusedFineStep
    from voxurf_b200.trainer import FINE_TRAIN
    world, n_rays, step = 2, 512, 15001
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, n_rays, step, True, ret, False, True, defer), nprocs=world, join=True)
    dev = torch.device('cuda', 0)
    m = _build(dev)
    fs = FusedFineStep(m, n_rays * world, FINE_TRAIN, RK, row_capacity=16384)
    bs = _batches(world, n_rays)
    cat = [torch.cat([b[i] for b in bs]).to(dev) for i in range(4)]
    for it in range(7):
        fs.step(*cat, step + it)
    refp = {'sdf': m.sdf.grid.detach().cpu(), 'k0': m.k0.grid.detach().contiguous().cpu(), 'mlp1': fs.mlp1.flat.detach().cpu()}
    for r in range(world):
        for k, lr in (('sdf', 5e-3), ('k0', 1e-1), ('mlp1', 3e-3)):
            d = (ret[r][1][k] - refp[k]).abs()
            assert float((d > 2e-2 * lr).float().mean()) < 1e-3 and float(d.max()) <= 8 * lr, (r, k, float(d.max()))
    assert torch.equal(ret[0][1]['sdf'], ret[1][1]['sdf']) and torch.equal(ret[0][1]['mlp1'], ret[1][1]['mlp1'])
    assert torch.equal(ret[0][1]['k0'], ret[1][1]['k0'])      # k0 ownership: the owner's update is stored into every replica
