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


def unpack(img, F, n_rows):
    """ACT row image [r/8][f/4][r%8][f%4] -> (n_rows, F)"""
    return img.view(-1, F // 4, 8, 4).permute(0, 2, 1, 3).reshape(-1, F)[:n_rows]


def pack(V):
    R, F = V.shape
    return V.view(R // 8, 8, F // 4, 4).permute(0, 2, 1, 3).contiguous().view(-1)


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


@pytest.mark.parametrize('R,FA,M_out,FB,N_in', [(64, 16, 16, 16, 16), (256, 192, 192, 192, 192), (1024, 8, 3, 192, 192),
                                                 (4096, 192, 192, 80, 80), (43008, 192, 192, 192, 192), (4096, 192, 192, 64, 54)])
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
