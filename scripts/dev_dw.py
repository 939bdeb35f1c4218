import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from voxurf_b200._lib import call
dev = 'cuda'
torch.manual_seed(0)


def pack(V, F):   # V (R, F) -> ACT image
    R = V.shape[0]
    return V.view(R // 8, 8, F // 4, 4).permute(0, 2, 1, 3).contiguous().view(-1)


def split(x):
    hi = (x.view(torch.int32) & -8192).view(torch.float32)
    return hi, x - hi


for (R, FA, M_out, FB, N_in) in [(64, 16, 16, 16, 16), (256, 192, 192, 192, 192), (1024, 8, 3, 192, 192), (4096, 192, 192, 80, 80), (43008, 192, 192, 192, 192)]:
    dY = torch.randn(R, FA, device=dev); H = torch.randn(R, FB, device=dev)
    dY[:, M_out:] = 0; H[:, N_in:] = 0
    C = torch.zeros(M_out, N_in, device=dev); cb = torch.zeros(M_out, device=dev)
    n = torch.tensor([R], dtype=torch.int32, device=dev)
    call('vx_mlp_dw', pack(dY, FA), FA, M_out, pack(H, FB), FB, N_in, n, R, C, C.stride(0), cb)
    torch.cuda.synchronize()
    ref = dY[:, :M_out].double().t() @ H[:, :N_in].double()
    rb = dY[:, :M_out].double().sum(0)
    print(R, FA, M_out, FB, N_in, 'dW rel', ((C.double() - ref).abs().max() / ref.abs().max()).item(), 'db rel', ((cb.double() - rb).abs().max() / rb.abs().max()).item())
    print('   |C| max', C.abs().max().item(), 'nonzero', (C != 0).sum().item(), 'cb', cb[:4].tolist())
