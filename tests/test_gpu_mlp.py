"""GPU: the tcgen05 (TF32x3) MLP kernels -- vx_mlp_prep(_batch), vx_mlp_chain, vx_mlp_dw -- against float64 torch.

The networks are the fine stage's rgbnet / k_rgbnet (lib/voxurf_fine.py:132-187: Linear+ReLU x3, Linear -> 3).
Tolerances (relative to the largest reference magnitude of the tensor, the accuracy class of an fp32 GEMM):
forward 5e-6, gradients 2e-5.  The backward reference uses the ReLU gates of the product's own forward activations:
a pre-activation within ~1e-6 of zero may fall on the other side of the gate than in the float64 forward (any fp32
GEMM has that property, cuBLAS included), which is a property of the inputs, not an error of the backward kernels.
"""
import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def make_net(d_in, width=192, depth=4):
    return nn.Sequential(nn.Linear(d_in, width), nn.ReLU(inplace=True),
                         *[nn.Sequential(nn.Linear(width, width), nn.ReLU(inplace=True)) for _ in range(depth - 2)],
                         nn.Linear(width, 3)).to(DEV)


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-300))


def act_index(R, F, device=DEV):
    """float index of element (r, f) in the ACT(F) row image (csrc/mlp_tc.cu act_offset): 4-row x 32-feature atoms,
    every row a 128-byte line whose 32-byte pieces are XOR-permuted with r % 4"""
    r = torch.arange(R, device=device).view(R, 1)
    f = torch.arange(F, device=device).view(1, F)
    return ((r // 4) * (F // 32) + f // 32) * 128 + (r % 4) * 32 + ((((f % 32) // 8) ^ (r % 4)) * 8) + f % 8


def unpack(img, F, n_rows):
    R = img.numel() // F
    return img[act_index(R, F, img.device).reshape(-1)].view(R, F)[:n_rows]


def pack(V):
    R, F = V.shape
    img = torch.empty(R * F, device=V.device)
    img[act_index(R, F, V.device).reshape(-1)] = V.reshape(-1)
    return img


@pytest.mark.parametrize('d_in,ld,cap,n_rows', [(79, 80, 2048, 1999), (54, 64, 2048, 1), (60, 64, 640, 640),
                                                (79, 80, 20096, 19000), (54, 64, 45056, 43001)])
def test_flat_mlp_forward_backward(d_in, ld, cap, n_rows):
    from voxurf_b200.mlp import FlatMLP
    torch.manual_seed(d_in + n_rows)
    net = make_net(d_in)
    f = FlatMLP(net, ld, d_in, tensor_core=True)
    f.alloc(cap)
    X = torch.zeros(cap, ld, device=DEV); X[:, :d_in] = torch.randn(cap, d_in, device=DEV)
    out = torch.full((cap, 3), 7.0, device=DEV); dX = torch.zeros(cap, ld, device=DEV)
    n_dev = torch.tensor([n_rows], dtype=torch.int32, device=DEV)
    d_out = torch.zeros(cap, 3, device=DEV); d_out[:n_rows] = torch.randn(n_rows, 3, device=DEV) * 1e-3
    f.forward(X, out, keep_activations=True, n_rows_dev=n_dev)
    f.backward(d_out, dX)
    torch.cuda.synchronize()

    W = [l.weight.detach().double() for l in f.linears]
    b = [l.bias.detach().double() for l in f.linears]
    W[0] = W[0][:, :d_in]
    acts = [X[:n_rows, :d_in].double()]
    for i in range(4):
        z = acts[-1] @ W[i].t() + b[i]
        acts.append(z.relu() if i < 3 else z)
    assert rel(out[:n_rows], acts[-1]) < 5e-6
    gates = []
    for i in range(3):
        H = unpack(f.H_img[i], 192, n_rows)
        assert rel(H, acts[i + 1]) < 5e-6
        gates.append(H > 0)
    # rows past n_rows: computed as zeros in the row images (they are operands of the weight-gradient GEMM)
    R = f.H_img[0].numel() // 192
    assert (unpack(f.H_img[2], 192, R)[n_rows:(n_rows + 127) // 128 * 128] == 0).all()
    dy = d_out[:n_rows].double()
    for i in range(3, -1, -1):
        dW, db = dy.t() @ acts[i], dy.sum(0)
        assert rel(f.linears[i].weight.grad[:, :acts[i].shape[1]], dW) < 2e-5, f'dW{i}'
        assert rel(f.linears[i].bias.grad, db) < 2e-5, f'db{i}'
        dy = dy @ W[i]
        if i > 0:
            dy = dy * gates[i - 1]
            assert rel(unpack(f.dH_img[i - 1], 192, n_rows), dy) < 2e-5, f'dH{i - 1}'
    assert rel(dX[:n_rows, :d_in], dy) < 2e-5
    assert (dX[n_rows:] == 0).all()


def test_flat_mlp_zero_rows_and_accumulation():
    """n_rows = 0 leaves the gradients untouched; a second backward accumulates (autograd .grad semantics)."""
    from voxurf_b200.mlp import FlatMLP
    torch.manual_seed(3)
    net = make_net(54)
    f = FlatMLP(net, 64, 54, tensor_core=True)
    f.alloc(256)
    X = torch.zeros(256, 64, device=DEV); X[:, :54] = torch.randn(256, 54, device=DEV)
    out, dX = torch.zeros(256, 3, device=DEV), torch.zeros(256, 64, device=DEV)
    d_out = torch.randn(256, 3, device=DEV)
    f.forward(X, out, keep_activations=True, n_rows_dev=torch.zeros(1, dtype=torch.int32, device=DEV))
    f.backward(d_out, dX)
    assert all((l.weight.grad == 0).all() and (l.bias.grad == 0).all() for l in f.linears) and (dX == 0).all()
    n_dev = torch.tensor([200], dtype=torch.int32, device=DEV)
    f.forward(X, out, keep_activations=True, n_rows_dev=n_dev)
    f.backward(d_out, dX)
    g1 = [l.weight.grad.clone() for l in f.linears]
    f.backward(d_out, dX)
    for l, g in zip(f.linears, g1):
        torch.testing.assert_close(l.weight.grad, 2 * g, rtol=1e-5, atol=1e-5 * float(g.abs().max()))


def test_flat_mlp_eval_forward_matches_cublas_path():
    """keep_activations=False (render): same outputs as the fp32 cuBLAS formulation of the same network."""
    from voxurf_b200.mlp import FlatMLP
    torch.manual_seed(4)
    net = make_net(79)
    X = torch.zeros(3000, 80, device=DEV); X[:, :79] = torch.randn(3000, 79, device=DEV)
    n_dev = torch.tensor([2900], dtype=torch.int32, device=DEV)
    o = []
    for tc in (True, False):
        f = FlatMLP(net, 80, 79, tensor_core=tc)
        out = torch.zeros(3000, 3, device=DEV)
        f.forward(X, out, keep_activations=False, n_rows_dev=n_dev if tc else None)
        o.append(out[:2900].clone())
    assert not torch.backends.cuda.matmul.allow_tf32, 'the cuBLAS comparison path must run in full fp32'
    assert rel(o[0], o[1].double()) < 5e-6


@pytest.mark.parametrize('R,FA,M_out,FB,N_in', [(64, 32, 16, 32, 16), (256, 192, 192, 192, 192), (1024, 32, 3, 192, 192),
                                                 (4096, 192, 192, 96, 80), (43008, 192, 192, 192, 192), (4096, 192, 192, 64, 54)])
def test_mlp_dw_direct(R, FA, M_out, FB, N_in):
    """vx_mlp_dw on packed row images: C += A^T B, c_bias += A^T 1 over the first *n_rows rows."""
    from voxurf_b200._lib import call
    torch.manual_seed(R + FA)
    dY = torch.randn(R, FA, device=DEV); H = torch.randn(R, FB, device=DEV)
    dY[:, M_out:] = 0; H[:, N_in:] = 0
    n_rows = R - 5 if R > 64 else R
    dY[n_rows:] = 0; H[n_rows:] = 0     # the chain kernel writes zeros past the row count
    C = torch.ones(M_out, N_in, device=DEV); cb = torch.ones(M_out, device=DEV)
    n = torch.tensor([n_rows], dtype=torch.int32, device=DEV)
    call('vx_mlp_dw', pack(dY), FA, M_out, pack(H), FB, N_in, n, R, C, C.stride(0), cb)
    ref = dY[:, :M_out].double().t() @ H[:, :N_in].double()
    rb = dY[:, :M_out].double().sum(0)
    assert rel(C - 1, ref) < 2e-5 and rel(cb - 1, rb) < 2e-5


def test_mlp_prep_images():
    """vx_mlp_prep(_batch): hi + lo == W to 2^-21 relative, hi has a 10-bit mantissa, zero padding, CH(Np) layout."""
    from voxurf_b200._lib import call
    torch.manual_seed(8)
    Wfull = torch.randn(100, 88, device=DEV)   # row stride 88, 79 valid columns
    W = Wfull[:, :79]
    for transpose in (0, 1):
        N, K = (79, 100) if transpose else (100, 79)
        Np, Kp = 128, 104
        hi = torch.full((Np * Kp,), 9.0, device=DEV); lo = torch.full((Np * Kp,), 9.0, device=DEV)
        call('vx_mlp_prep', Wfull, N, K, Wfull.stride(0), Np, Kp, transpose, hi, lo)
        un = lambda im: im.view(Kp // 4, Np, 4).permute(1, 0, 2).reshape(Np, Kp)
        L = W.t() if transpose else W
        h, l = un(hi), un(lo)
        assert (h[N:] == 0).all() and (h[:, K:] == 0).all() and (l[N:] == 0).all() and (l[:, K:] == 0).all()
        assert ((h[:N, :K].view(torch.int32) & 0x1fff) == 0).all()
        assert ((h + l)[:N, :K] - L).abs().max() <= 2.0 ** -21 * L.abs().max()


def test_mlp_dw_batch_matches_single_jobs():
    """vx_mlp_dw_batch (8 GEMMs of different shapes in one launch, SMs dealt by cost) == the GEMMs one by one."""
    from voxurf_b200._lib import call
    torch.manual_seed(11)
    R, n_rows = 6144, 6001
    shapes = [(32, 3, 192, 192), (192, 192, 192, 192), (192, 192, 192, 192), (192, 192, 96, 79),
              (32, 3, 192, 192), (192, 192, 192, 192), (192, 192, 192, 192), (192, 192, 64, 54)]
    n = torch.tensor([n_rows], dtype=torch.int32, device=DEV)
    ptrs, dims, keep, refs = [], [], [], []
    for FA, M_out, FB, N_in in shapes:
        dY = torch.randn(R, FA, device=DEV); H = torch.randn(R, FB, device=DEV)
        dY[:, M_out:] = 0; H[:, N_in:] = 0; dY[n_rows:] = 0; H[n_rows:] = 0
        A, B = pack(dY), pack(H)
        C = torch.zeros(M_out, N_in, device=DEV); cb = torch.zeros(M_out, device=DEV)
        keep.append((A, B, C, cb))
        ptrs += [A.data_ptr(), B.data_ptr(), C.data_ptr(), cb.data_ptr()]
        dims += [FA, M_out, FB, N_in, C.stride(0)]
        refs.append((dY[:, :M_out].double().t() @ H[:, :N_in].double(), dY[:, :M_out].double().sum(0)))
    call('vx_mlp_dw_batch', len(shapes), ptrs, dims, n, R)
    for (A, B, C, cb), (rC, rb) in zip(keep, refs):
        assert rel(C, rC) < 2e-5 and rel(cb, rb) < 2e-5


def test_umma_mn_major_tf32_operand_layout():
    """Pins the operand layout the row images rely on: one M = 128, K = 8 TF32 MMA (vx_umma_probe) with A and B read
    MN-major through layout type 1 from images built with act_index()."""
    from voxurf_b200._lib import call
    torch.manual_seed(2)
    N = 64
    A = torch.randint(-8, 9, (8, 128)).float()     # [k][m]: exactly representable in TF32
    B = torch.randint(-8, 9, (8, N)).float()       # [k][n]
    a_img = torch.zeros(8192, device=DEV); b_img = torch.zeros(8192, device=DEV)
    a_img[:8 * 128] = pack(A.to(DEV)); b_img[:8 * N] = pack(B.to(DEV))
    desc = lambda lbo, sbo: ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (1 << 61)
    idesc = (1 << 4) | (2 << 7) | (2 << 10) | (1 << 15) | (1 << 16) | ((N >> 3) << 17) | ((128 >> 4) << 24)
    D = torch.zeros(128, N, device=DEV)
    call('vx_umma_probe', a_img, 8192, b_img, 8192, desc(512, 16 * 128), desc(512, 16 * N), idesc, N, D)
    assert torch.equal(D.cpu(), A.t() @ B)
