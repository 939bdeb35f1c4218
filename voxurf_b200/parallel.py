"""Data parallelism over rays.  The reference is single-GPU (no distributed code at all); this is
the one exchange step the path needs when rays are sharded: grids and MLPs are replicated, every rank renders its
own ray batch, and the gradients of the batch-mean loss are averaged across ranks before the TV / Adam step
(run.py:604 is a mean over the batch, so the global-batch gradient is the mean of the per-rank gradients).

One process per GPU, torch.distributed (NCCL over NVLink/NVSwitch; gloo in the CPU tests).  Small tensors (MLP
weights) are coalesced into one flat bucket so the exchange is three collectives per step: k0, sdf, MLP bucket.
Inference and mesh queries do not exchange anything: `shard_range` splits chunks / voxel slabs.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous [lo, hi) share of n_items for `rank` (image chunks at render time, X-planes of a mesh lattice)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dense(t):
    if t.is_contiguous():
        return t
    if t.dim() == 5 and t.is_contiguous(memory_format=torch.channels_last_3d):
        return t.permute(0, 2, 3, 4, 1)
    raise RuntimeError('gradient must be dense')


class GradSync:
    """Averages .grad of every trainable parameter across ranks, in place."""

    def __init__(self, model, world, small_numel=1 << 20, group=None, tensors=None):
        """tensors: explicit list of tensors whose .grad is exchanged one collective each (the fused step passes its
        grids and its flat per-network MLP buffers); default: every trainable parameter of `model`, small ones bucketed."""
        self.world, self.group = world, group
        if tensors is not None:
            self.large, self.small = list(tensors), []
        else:
            params = [p for p in model.parameters() if p.requires_grad]
            self.large = [p for p in params if p.numel() >= small_numel]
            self.small = [p for p in params if p.numel() < small_numel]
        self._flat = None

    def __call__(self, model=None):
        if self.world <= 1:
            return
        for p in self.large:
            if p.grad is not None:
                g = _dense(p.grad)
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
                g.div_(self.world)
        small = [p for p in self.small if p.grad is not None]
        if small:
            n = sum(p.numel() for p in small)
            if self._flat is None or self._flat.numel() != n or self._flat.device != small[0].grad.device:
                self._flat = torch.empty(n, dtype=torch.float32, device=small[0].grad.device)
            o = 0
            for p in small:
                self._flat[o:o + p.numel()].copy_(p.grad.reshape(-1))
                o += p.numel()
            dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.group)
            self._flat.div_(self.world)
            o = 0
            for p in small:
                p.grad.copy_(self._flat[o:o + p.numel()].view_as(p.grad))
                o += p.numel()
