"""2-GPU debugging of the k0 ownership path (torchrun --nproc-per-node 2 scripts/debug_k0_peers.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['CUDA_LAUNCH_BLOCKING'] = '1'
import torch
import torch.distributed as dist
from voxurf_b200._lib import call

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dev = torch.device('cuda', rank)
dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from tests.test_gpu_multi import _build, _batches, RK
from voxurf_b200.fused import FusedFineStep
from voxurf_b200.trainer import FINE_TRAIN


def say(*a):
    torch.cuda.synchronize()
    print(f'[{rank}]', *a, flush=True)


m = _build(dev)
fs = FusedFineStep(m, 512, FINE_TRAIN, RK, row_capacity=8192, world=world, rank=rank)
say('k0_owned', fs.k0_owned, fs.k0_peer_note, 'slab', fs.slab)
say('peer ptrs', [hex(p) for p in getattr(fs, '_k0_peer_ptrs', [])])
b = [t.to(dev) for t in _batches(world, 512)[rank]]
fs.forward_backward(*b, 15003)
say('fwd/bwd ok')
fs._sync_begin()
say('sync_begin ok')
fs._sync_k0()
say('sync_k0 ok')
fs._sync_end()
say('sync_end ok')
fs.regularise(15003)
fs.optimizer_step(only=('k0',))
say('k0 adam ok')
fs.optimizer_step(only=('sdf', 'rgbnet', 'k_rgbnet'))
fs.sync_params()
say('all ok', float(m.k0.grid.double().sum()))
fs.shutdown()
dist.destroy_process_group()
