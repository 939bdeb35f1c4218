"""Time vx_mlp_dw_batch alone on the bench shapes (development)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from voxurf_b200._lib import call
dev = 'cuda'
R, n_rows = 45056, int(os.environ.get('NROWS', '36400'))
shapes = [(32, 3, 192, 192), (192, 192, 192, 192), (192, 192, 192, 192), (192, 192, 96, 79),
          (32, 3, 192, 192), (192, 192, 192, 192), (192, 192, 192, 192), (192, 192, 64, 54)]
n = torch.tensor([n_rows], dtype=torch.int32, device=dev)
ptrs, dims, keep = [], [], []
for FA, M_out, FB, N_in in shapes:
    A = torch.randn(R * FA, device=dev); B = torch.randn(R * FB, device=dev)
    C = torch.zeros(M_out, N_in, device=dev); cb = torch.zeros(M_out, device=dev)
    keep.append((A, B, C, cb))
    ptrs += [A.data_ptr(), B.data_ptr(), C.data_ptr(), cb.data_ptr()]
    dims += [FA, M_out, FB, N_in, C.stride(0)]
sel = [int(x) for x in os.environ.get('JOBS', '0,1,2,3,4,5,6,7').split(',')]
p = sum([ptrs[4 * j:4 * j + 4] for j in sel], []); d = sum([dims[5 * j:5 * j + 5] for j in sel], [])
for _ in range(3):
    call('vx_mlp_dw_batch', len(sel), p, d, n, R)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    call('vx_mlp_dw_batch', len(sel), p, d, n, R)
e1.record(); torch.cuda.synchronize()
print('rows', n_rows, 'experiment', os.environ.get('VX_DW_EXPERIMENT', '0'), 'jobs', sel, 'us per launch %.1f' % (e0.elapsed_time(e1) / 20 * 1000))
