import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from voxurf_b200.mlp import FlatMLP
dev = 'cuda'


def mk(d_in, width=192, depth=4):
    return nn.Sequential(nn.Linear(d_in, width), nn.ReLU(inplace=True),
                         *[nn.Sequential(nn.Linear(width, width), nn.ReLU(inplace=True)) for _ in range(depth - 2)],
                         nn.Linear(width, 3)).to(dev)


rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max()).item()
for d_in, ld, cap, n_rows in [(79, 80, 2048, 1999), (54, 64, 2048, 1999), (79, 80, 20000, 19000), (79, 80, 61440, 43001), (54, 64, 40000, 30000)]:
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
    # reference with intermediates
    lins = [m for m in ref.modules() if isinstance(m, nn.Linear)]
    h = X[:n_rows, :d_in].double().requires_grad_(True)
    acts = [h]
    for i, l in enumerate(lins):
        z = acts[-1] @ l.weight.t() + l.bias
        if i < 3:
            z = z.relu()
        z.retain_grad()
        acts.append(z)
    acts[-1].backward(d_out[:n_rows].double())
    msg = f'd_in {d_in} cap {cap} rows {n_rows}: fwd {rel(out[:n_rows], acts[-1]):.1e}'
    un = lambda im, F: im.view(-1, F // 4, 8, 4).permute(0, 2, 1, 3).reshape(-1, F)[:n_rows]
    for i in range(3):
        msg += f' H{i} {rel(un(f.H_img[i], 192), acts[i + 1]):.1e} dH{i} {rel(un(f.dH_img[i], 192), acts[i + 1].grad * (acts[i + 1] > 0)):.1e}'
    msg += f' dX {rel(dX[:n_rows, :d_in], h.grad):.1e}'
    for l, lr in zip(f.linears, lins):
        msg += f' dW {rel(l.weight.grad, lr.weight.grad):.1e}'
    print(msg)
