"""The two colour MLPs of the fine model (rgbnet 79->192->192->192->3, k_rgbnet 54->...; lib/voxurf_fine.py:132-187)
run over the fixed-capacity row matrices of the fused step with explicit forward / backward passes and flat
parameter + gradient storage (one Adam launch per network, no autograd graph).

This is the one dense-contraction stage of the path.  Default: the hand-written tcgen05 kernels of
csrc/mlp_tc.cu (TF32x3 split, fp32-grade accuracy): one fused launch per layer chain (forward, and the dX chain of
the backward pass) that keeps the activations on chip, and one batched launch for all weight / bias gradients.
tensor_core=False keeps the plain fp32 cuBLAS formulation (torch.addmm / mm on preallocated buffers, TF32 off) as a
comparison path for the tests.
"""
import torch
import torch.nn as nn

CHAIN_TIMINGS = None   # bench.py: list collecting ((start, end) CUDA events, flops per row) of every vx_mlp_chain launch


class FlatMLP:
    def __init__(self, seq, ld_in, d_in, tensor_core=True):
        """seq: nn.Sequential of Linear / ReLU (possibly nested).  The first layer's weight is stored padded to
        `ld_in` input columns (zeros) so the padded row matrix is consumed without a strided copy."""
        self.linears = [m for m in seq.modules() if isinstance(m, nn.Linear)]
        assert self.linears[0].in_features == d_in
        dev = self.linears[0].weight.device
        self.ld_in, self.d_in = ld_in, d_in
        shapes = []
        for i, l in enumerate(self.linears):
            shapes.append((l.out_features, ld_in if i == 0 else l.in_features))
        n = sum(o * i + o for o, i in shapes)
        n = (n + 3) // 4 * 4
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat.grad = torch.zeros_like(self.flat)
        self.W, self.b, self.dW, self.db = [], [], [], []
        o = 0
        for (out_f, in_f), l in zip(shapes, self.linears):
            W = self.flat[o:o + out_f * in_f].view(out_f, in_f)
            dW = self.flat.grad[o:o + out_f * in_f].view(out_f, in_f)
            o += out_f * in_f
            b = self.flat[o:o + out_f]
            db = self.flat.grad[o:o + out_f]
            o += out_f
            W[:, :l.in_features].copy_(l.weight.data)
            b.copy_(l.bias.data)
            # the module's parameters become views of the flat storage: state_dict / external code see live values
            l.weight.data = W[:, :l.in_features]
            l.bias.data = b
            l.weight.grad = dW[:, :l.in_features]
            l.bias.grad = db
            self.W.append(W); self.b.append(b); self.dW.append(dW); self.db.append(db)
        self.H = None
        self.tensor_core = tensor_core
        self.tc_fwd = self.tc_bwd = None
        if tensor_core:
            n = len(self.W)
            # forward chain: the hidden layers on the tensor cores, the tiny last layer (N = 3) in exact fp32 on the CUDA
            # cores inside the last hidden layer's epilogue (csrc/mlp_tc.cu)
            self.tc_fwd = TensorCoreChain([dict(W=self.W[i], bias=self.b[i], relu=True) for i in range(n - 1)],
                                          final=(self.W[-1], self.b[-1]))
            # dX chain: dH_{l-1} = (dH_l W_l) * (H_{l-1} > 0), i.e. the same kernel on the transposed weights, last layer first
            self.tc_bwd = TensorCoreChain([dict(W=self.W[i], bias=None, relu=False) for i in range(n - 1, -1, -1)], transpose=True)

    def rebind_grad(self, buf):
        """Move the flat gradient storage into `buf` (numel == flat.numel()): lets several networks share one buffer so that
        a data-parallel exchange is ONE collective for all of them."""
        assert buf.numel() == self.flat.numel() and buf.dtype == torch.float32
        buf.copy_(self.flat.grad)
        self.flat.grad = buf
        o = 0
        for i, l in enumerate(self.linears):
            out_f, in_f = self.W[i].shape
            self.dW[i] = buf[o:o + out_f * in_f].view(out_f, in_f)
            o += out_f * in_f
            self.db[i] = buf[o:o + out_f]
            o += out_f
            l.weight.grad = self.dW[i][:, :l.in_features]
            l.bias.grad = self.db[i]

    def alloc(self, cap):
        dev = self.flat.device
        self.cap = cap
        if self.tensor_core:
            # raw fp32 "row images" of every activation and activation gradient in the MN-major UMMA operand layout
            # (csrc/mlp_tc.cu, act_offset; feature count padded to 32): written by the chain epilogues, consumed by the
            # split-K weight-gradient GEMM with plain bulk copies
            self.R = (cap + 127) // 128 * 128
            img = lambda F: torch.zeros(self.R * F, dtype=torch.float32, device=dev)
            assert all(l.out_features % 32 == 0 for l in self.linears[:-1]), 'hidden widths must be multiples of 32'
            self.F_in = _pad(self.tc_fwd.Kp[0], 32)
            self.F_out = _pad(self.tc_bwd.Kp[0], 32)
            self.done = torch.zeros(2 * (self.R // 128), dtype=torch.int32, device=dev)   # vx_mlp_chain_batch tile flags
            self.X_img = img(self.F_in)
            self.H_img = [img(l.out_features) for l in self.linears[:-1]]
            self.dH_img = [img(l.out_features) for l in self.linears[:-1]]
            # ReLU gates of the hidden activations, one bit per (row, feature): written by the forward chain, read by the dX chain
            self.G_img = [torch.zeros(self.R * 4, dtype=torch.int64, device=dev) for l in self.linears[:-1]]
            self.dY_img = img(self.F_out)
            self.H = self.dH = []
            return
        self.H = [torch.empty(cap, l.out_features, dtype=torch.float32, device=dev) for l in self.linears[:-1]]
        self.dH = [torch.empty_like(h) for h in self.H]

    def chains(self, train=True):
        return [self.tc_fwd, self.tc_bwd] if train else [self.tc_fwd]

    def forward(self, X, out, keep_activations=True, n_rows_dev=None, prepared=False):
        """X (cap, ld_in) -> out (cap, 3).  Hidden activations stay in self.H for the backward pass.
        prepared: the caller already refreshed the weight images of self.chains(keep_activations) (prepare_chains)."""
        if self.H is None or getattr(self, 'cap', None) != X.shape[0]:
            self.alloc(X.shape[0])
        self._X = X
        if self.tensor_core:
            assert n_rows_dev is not None
            self._n = n_rows_dev
            if not prepared:   # the optimizer changed the weights since the last step
                prepare_chains(self.chains(keep_activations))
            run_chain_jobs([self.forward_job(X, out, keep_activations)], n_rows_dev, X.shape[0], None)
            return out
        h = X
        for i in range(len(self.linears) - 1):
            torch.addmm(self.b[i], h, self.W[i].t(), out=self.H[i])
            self.H[i].relu_()
            h = self.H[i]
        torch.addmm(self.b[-1], h, self.W[-1].t(), out=out)
        self._X = X
        return out

    def forward_job(self, X, out, keep_activations=True, patch=None):
        """Job descriptor of this network's forward chain for run_chain_jobs.  patch = (tensor (cap, ld), col0, n, dep):
        input columns [col0, col0 + n) are read from `tensor` (written by job `dep` of the same launch) instead of X."""
        self._X = X
        if keep_activations:
            return self.tc_fwd.job(X, self.ld_in, out, out.shape[1], imgs=self.H_img, x_img=self.X_img, patch=patch, gates_out=self.G_img)
        return self.tc_fwd.job(X, self.ld_in, out, out.shape[1], patch=patch)

    def backward_job(self, d_out, dX):
        """Job descriptor of the dX chain: chain layer j <-> network layer n-1-j, ReLU gates from the forward row images."""
        n = len(self.linears)
        rev = list(range(n - 2, -1, -1))
        return self.tc_bwd.job(d_out, d_out.shape[1], dX, dX.shape[1], imgs=[self.dH_img[i] for i in rev],
                               masks=[self.G_img[i] for i in rev], x_img=self.dY_img)

    def dw_jobs(self):
        """(ptrs, dims) of this network's weight-gradient GEMMs for vx_mlp_dw_batch: dW_i = dY_i^T H_{i-1}, db_i = dY_i^T 1
        on the row images left by forward() / backward()."""
        n = len(self.linears)
        ptrs, dims = [], []
        for i in range(n):
            A, FA = (self.dY_img, self.F_out) if i == n - 1 else (self.dH_img[i], self.linears[i].out_features)
            B, FB = (self.X_img, self.F_in) if i == 0 else (self.H_img[i - 1], self.linears[i - 1].out_features)
            ptrs += [A.data_ptr(), B.data_ptr(), self.dW[i].data_ptr(), self.db[i].data_ptr()]
            dims += [FA, self.linears[i].out_features, FB, self.dW[i].shape[1], self.dW[i].stride(0)]
        return ptrs, dims

    def backward(self, d_out, dX, defer_dw=False):
        """d_out (cap, 3) -> dX (cap, ld_in); accumulates weight / bias gradients into the flat gradient buffer.
        defer_dw: only run the dX chain; the caller batches the weight-gradient GEMMs of several networks into one
        launch with run_dw_batch()."""
        n = len(self.linears)
        if self.tensor_core:
            # dX chain (one launch): chain layer j <-> network layer n-1-j; hidden gradients land feature-major in dHT
            run_chain_jobs([self.backward_job(d_out, dX)], self._n, d_out.shape[0], None)
            if not defer_dw:
                run_dw_batch([self])
            return dX
        dy = d_out
        for i in range(n - 1, -1, -1):
            h_in = self.H[i - 1] if i > 0 else self._X
            self.db[i].add_(dy.sum(0))
            self.dW[i].addmm_(dy.t(), h_in)
            if i > 0:
                torch.mm(dy, self.W[i], out=self.dH[i - 1])
                self.dH[i - 1].masked_fill_(self.H[i - 1] <= 0, 0.0)
                dy = self.dH[i - 1]
            else:
                torch.mm(dy, self.W[0], out=dX)
        return dX


# ------------------------------------------------------------------------------------------------
# tensor-core path (tcgen05, TF32x3): csrc/mlp_tc.cu
# ------------------------------------------------------------------------------------------------
def _pad(v, m):
    return (v + m - 1) // m * m


def run_dw_batch(mlps, fx=None):
    """Weight / bias gradients of several FlatMLPs (same row count and capacity) with one launch per 8 layers.
    fx = (int64 accumulators, gradient buffer they mirror, scale): deterministic fixed-point accumulation."""
    from ._lib import call
    ptrs, dims = [], []
    for m in mlps:
        assert m._n is mlps[0]._n and m.cap == mlps[0].cap
        p, d = m.dw_jobs()
        ptrs += p; dims += d
    for j in range(0, len(ptrs) // 4, 8):
        n = min(8, len(ptrs) // 4 - j)
        if fx is not None:
            call('vx_mlp_dw_batch_fx', n, ptrs[4 * j:4 * (j + n)], dims[5 * j:5 * (j + n)], mlps[0]._n, mlps[0].cap, fx[0], fx[1], fx[2])
        else:
            call('vx_mlp_dw_batch', n, ptrs[4 * j:4 * (j + n)], dims[5 * j:5 * (j + n)], mlps[0]._n, mlps[0].cap)


JOB_STRIDE, PTR_STRIDE = 26, 30   # csrc/mlp_tc.cu MC_JOB_STRIDE (dims per job), MC_PTR_STRIDE (pointers per job)


def run_chain_jobs(jobs, n_rows_dev, capacity, done_flags):
    """Up to two layer chains (e.g. both networks' forward chains, or both dX chains) in ONE vx_mlp_chain_batch launch:
    their 128-row tiles share the SMs, so 2 x 353 tiles take 4.8 waves instead of 2 x 3."""
    from ._lib import call
    ptrs, dims = [], []
    for p, d in jobs:
        ptrs += p; dims += d
    timed = CHAIN_TIMINGS is not None
    if timed:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    call('vx_mlp_chain_batch', len(jobs), ptrs, dims, n_rows_dev, capacity, done_flags)
    if timed:
        ev[1].record()
        CHAIN_TIMINGS.append((ev, sum(d.flops for _, d in jobs)))


def prepare_chains(chains):
    """(Re)write the hi/lo TF32 weight images of several chains with one launch per 16 layers (vx_mlp_prep_batch)."""
    from ._lib import call
    ptrs, dims = [], []
    for c in chains:
        p, d = c.prep_jobs()
        ptrs += p; dims += d
    for j in range(0, len(ptrs) // 3, 16):
        n = min(16, len(ptrs) // 3 - j)
        call('vx_mlp_prep_batch', n, ptrs[3 * j:3 * (j + n)], dims[6 * j:6 * (j + n)])


class _Dims(list):
    """dims list of a chain job that also remembers the chain's flops per row (bench.py roofline bookkeeping)"""

    def __init__(self, v, flops):
        super().__init__(v)
        self.flops = flops


class TensorCoreChain:
    """A chain of up to 4 dense layers executed by one fused tcgen05 kernel launch (vx_mlp_chain).

    layers: list of dicts {W: (N,K) tensor view (row stride ldw), bias: (N,) or None, relu: bool}.  `prepare()` must
    be called whenever the weights changed (it writes the hi/lo TF32 split images, zero-padded)."""

    def __init__(self, layers, transpose=False, final=None):
        """final = (W (n_out, K) view, bias (n_out,)): an extra last layer computed on the CUDA cores (vx_mlp_chain_batch)."""
        from ._lib import call
        self._call = call
        self.layers = layers
        self.transpose = transpose
        self.final = final
        dev = layers[0]['W'].device
        self.Kp, self.Np, self.N, self.K = [], [], [], []
        n = len(layers)
        for i, L in enumerate(layers):
            N, K = (L['W'].shape[1], L['W'].shape[0]) if transpose else (L['W'].shape[0], L['W'].shape[1])
            self.N.append(N); self.K.append(K)
            self.Kp.append(_pad(K, 8))
            self.Np.append(_pad(N, 32) if i + 1 < n else _pad(N, 16))
        if final is not None:
            self.Np[-1] = _pad(self.N[-1], 32)
        for i in range(n - 1):
            assert self.Np[i] == self.Kp[i + 1], 'hidden widths must be multiples of 32 and chain'
        self.W_hi = [torch.zeros(self.Np[i], self.Kp[i], dtype=torch.float32, device=dev) for i in range(n)]
        self.W_lo = [torch.zeros_like(w) for w in self.W_hi]

    def prep_jobs(self):
        ptrs, dims = [], []
        for i, L in enumerate(self.layers):
            W = L['W']
            assert W.stride(1) == 1
            ptrs += [W.data_ptr(), self.W_hi[i].data_ptr(), self.W_lo[i].data_ptr()]
            dims += [self.N[i], self.K[i], W.stride(0), self.Np[i], self.Kp[i], int(self.transpose)]
        return ptrs, dims

    def prepare(self):
        prepare_chains([self])

    def flops_per_row(self):
        f = 2 * sum(k * n_ for k, n_ in zip(self.K, self.N))
        if self.final is not None:
            f += 2 * self.final[0].shape[0] * self.final[0].shape[1]
        return f

    def job(self, X, k0, Y, n_out, imgs=None, masks=None, x_img=None, patch=None, gates_out=None):
        """(ptrs, dims) of this chain as a vx_mlp_chain_batch job (see csrc/mlp_tc.cu for the packing)."""
        pick = lambda lst, i: lst[i] if (lst is not None and i < len(lst)) else None
        addr = lambda t: t.data_ptr() if t is not None else 0
        Wf, bf = self.final if self.final is not None else (None, None)
        pt, pcol, pn, dep = patch if patch is not None else (None, 0, 0, -1)
        ptrs = [X.data_ptr(), addr(x_img), Y.data_ptr(), addr(Wf), addr(bf), addr(pt)]
        dims = [X.stride(0), k0, len(self.layers), Y.stride(0), n_out, Wf.stride(0) if Wf is not None else 0, pcol, pn,
                pt.stride(0) if pt is not None else 0, dep]
        for i, L in enumerate(self.layers):
            ptrs += [self.W_hi[i].data_ptr(), self.W_lo[i].data_ptr(), addr(L.get('bias')), addr(pick(imgs, i)), addr(pick(masks, i)),
                     addr(pick(gates_out, i))]
            dims += [self.Kp[i], self.Np[i], self.N[i], int(bool(L.get('relu', False)))]
        ptrs += [0] * (PTR_STRIDE - len(ptrs))
        dims += [0] * (JOB_STRIDE - len(dims))
        return ptrs, _Dims(dims, self.flops_per_row())

    def run(self, X, k0, n_rows_dev, Y, n_out, imgs=None, masks=None, x_img=None):
        """X (cap, ldx) with k0 valid columns -> Y (cap, ldy)[:, :n_out].  imgs[l] = (hi, lo) CH(Np) images written for
        layer l's output (ACT row image), masks[l] = ReLU-gate words (int64 x 4 per row, csrc/mlp_tc.cu MlpLayer.gate) applied to layer l's output (ReLU backward), x_img = row image of X."""
        if self.final is not None:
            return run_chain_jobs([self.job(X, k0, Y, n_out, imgs, masks, x_img)], n_rows_dev, X.shape[0], None)
        n = len(self.layers)
        ptrs, dims = [], []
        pick = lambda lst, i: lst[i] if (lst is not None and i < len(lst)) else None
        addr = lambda t: t.data_ptr() if t is not None else 0
        for i, L in enumerate(self.layers):
            ptrs += [self.W_hi[i].data_ptr(), self.W_lo[i].data_ptr(), addr(L.get('bias')), addr(pick(imgs, i)),
                     addr(pick(masks, i))]
            dims += [self.Kp[i], self.Np[i], self.N[i], int(bool(L.get('relu', False)))]
        timed = CHAIN_TIMINGS is not None
        if timed:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        self._call('vx_mlp_chain', X, X.stride(0), k0, n_rows_dev, X.shape[0], n, ptrs, dims, Y, Y.stride(0), n_out,
                   x_img)
        if timed:
            ev[1].record()
            CHAIN_TIMINGS.append((ev, 2 * sum(k * nn_ for k, nn_ in zip(self.K, self.N))))   # flops per row
