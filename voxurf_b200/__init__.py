"""voxurf_b200: B200-native (sm_100a) implementation of Voxurf's ray-batch volume-rendering hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); every operator of the path
is a hand-written CUDA kernel reached through the C-ABI library `libvoxurf_b200.so`
(include/voxurf_b200.h).  There is no CPU fallback: calling an operator without the built library
or without a CUDA device raises.
"""
__version__ = '0.1.0'
