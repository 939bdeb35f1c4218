"""Dev check of the tcgen05 MLP chain against float64 torch (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from voxurf_b200.mlp import TensorCoreChain

torch.manual_seed(0)
dev = 'cuda'


def run_case(dims, cap, n_rows, k0_valid=None, relu=True):
    Ws = [torch.randn(dims[i + 1], dims[i], device=dev) / dims[i] ** 0.5 for i in range(len(dims) - 1)]
    bs = [torch.randn(dims[i + 1], device=dev) * 0.1 for i in range(len(dims) - 1)]
    layers = [dict(W=W, bias=b, relu=(relu and i + 1 < len(Ws))) for i, (W, b) in enumerate(zip(Ws, bs))]
    ch = TensorCoreChain(layers)
    ch.prepare()
    ldx = (dims[0] + 15) // 16 * 16
    X = torch.zeros(cap, ldx, device=dev)
    X[:, :dims[0]] = torch.randn(cap, dims[0], device=dev)
    Y = torch.full((cap, dims[-1]), float('nan'), device=dev)
    R = (cap + 127) // 128 * 128
    imgs = [torch.zeros(R * dims[i + 1], device=dev) for i in range(len(dims) - 2)]
    n_dev = torch.tensor([n_rows], dtype=torch.int32, device=dev)
    ch.run(X, dims[0], n_dev, Y, dims[-1], imgs=imgs)
    torch.cuda.synchronize()
    H = [a.view(R // 8, dims[i + 1] // 4, 8, 4).permute(0, 2, 1, 3).reshape(R, dims[i + 1]) for i, a in enumerate(imgs)]
    h = X[:n_rows, :dims[0]].double()
    ok = True
    for i, (W, b) in enumerate(zip(Ws, bs)):
        h = h @ W.double().t() + b.double()
        if i + 1 < len(Ws):
            if relu:
                h = h.relu()
            err = ((H[i][:n_rows].double() - h).abs().max() / h.abs().max()).item()
            print(f'   hidden {i}: rel err {err:.3e}')
            ok &= err < 3e-6
    err = ((Y[:n_rows].double() - h).abs().max() / h.abs().max()).item()
    untouched = torch.isnan(Y[n_rows:]).all().item() if n_rows < cap else True
    print(f'dims {dims} cap {cap} rows {n_rows}: out rel err {err:.3e} tail untouched {untouched}')
    return ok and err < 3e-6 and untouched


ok = True
ok &= run_case([192, 16], 128, 128)                      # one K=192 layer, N=16
ok &= run_case([64, 192, 3], 256, 200)
ok &= run_case([80, 192, 192, 192, 3], 1024, 1000)
ok &= run_case([64, 192, 192, 192, 3], 70000, 61234)
print('ALL OK' if ok else 'FAILED')
import time
# timing of the full-size chain
dims = [80, 192, 192, 192, 3]
Ws = [torch.randn(dims[i + 1], dims[i], device=dev) / dims[i] ** 0.5 for i in range(4)]
bs = [torch.randn(dims[i + 1], device=dev) * 0.1 for i in range(4)]
ch = TensorCoreChain([dict(W=W, bias=b, relu=i < 3) for i, (W, b) in enumerate(zip(Ws, bs))]); ch.prepare()
cap = 61440
X = torch.randn(cap, 80, device=dev); Y = torch.empty(cap, 3, device=dev)
H = [torch.zeros(cap * 192, device=dev) for _ in range(3)]
n_dev = torch.tensor([43000], dtype=torch.int32, device=dev)
for _ in range(3):
    ch.run(X, 80, n_dev, Y, 3, imgs=H)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ch.run(X, 80, n_dev, Y, 3, imgs=H)
e1.record(); torch.cuda.synchronize()
print('chain fwd 43000 rows: %.1f us' % (e0.elapsed_time(e1) / 20 * 1e3))


# ------------------------------------------------------------------ full fwd + bwd through FlatMLP (tensor-core path)
import torch.nn as nn
from voxurf_b200.mlp import FlatMLP


def mk(d_in, width=192, depth=4):
    return nn.Sequential(nn.Linear(d_in, width), nn.ReLU(inplace=True),
                         *[nn.Sequential(nn.Linear(width, width), nn.ReLU(inplace=True)) for _ in range(depth - 2)],
                         nn.Linear(width, 3)).to(dev)


for d_in, ld, cap, n_rows in [(79, 80, 2048, 1999), (54, 64, 61440, 43001)]:
    torch.manual_seed(1)
    net = mk(d_in)
    ref = mk(d_in).double()
    ref.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
    f = FlatMLP(net, ld, d_in, tensor_core=True)
    f.alloc(cap)
    X = torch.zeros(cap, ld, device=dev); X[:, :d_in] = torch.randn(cap, d_in, device=dev)
    out = torch.zeros(cap, 3, device=dev); dX = torch.zeros(cap, ld, device=dev)
    n_dev = torch.tensor([n_rows], dtype=torch.int32, device=dev)
    d_out = torch.zeros(cap, 3, device=dev); d_out[:n_rows] = torch.randn(n_rows, 3, device=dev) * 1e-3
    f.forward(X, out, keep_activations=True, n_rows_dev=n_dev)
    f.backward(d_out, dX)
    torch.cuda.synchronize()
    Xr = X[:n_rows, :d_in].double().requires_grad_(True)
    yr = ref(Xr)
    yr.backward(d_out[:n_rows].double())
    rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max()).item()
    print(f'net {d_in}: fwd {rel(out[:n_rows], yr):.2e}  dX {rel(dX[:n_rows, :d_in], Xr.grad):.2e}', end='')
    for l, lr in zip(f.linears, [m for m in ref.modules() if isinstance(m, nn.Linear)]):
        print(f'  dW {rel(l.weight.grad, lr.weight.grad):.2e} db {rel(l.bias.grad, lr.bias.grad):.2e}', end='')
    print()
    for _ in range(3):
        f.forward(X, out, True, n_dev); f.backward(d_out, dX)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for _ in range(10):
        f.forward(X, out, True, n_dev)
    e1.record()
    for _ in range(10):
        f.backward(d_out, dX)
    e2.record(); torch.cuda.synchronize()
    print(f'   fwd {e0.elapsed_time(e1) * 100:.0f} us, bwd {e1.elapsed_time(e2) * 100:.0f} us  ({n_rows} rows)')
