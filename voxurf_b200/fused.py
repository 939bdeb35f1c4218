"""Sync-free fused fine-stage step: the B200-first execution of Voxurf.forward + losses + backward + TV +
per-voxel Adam (lib/voxurf_fine.py:620-802, run.py:600-659) on persistent device buffers.

What this replaces, per training iteration of the reference: ~150 kernel launches, >= 13 host syncs, ~10
boolean-mask compactions and a dense autograd graph.  Here the iteration is ~25 kernels (+3 on TV iterations), replayed
as ONE CUDA graph (use_graph=True; at N > 1 with its collectives captured), ZERO host syncs and no per-step allocation:
every data-dependent count (M2 = samples after bbox + mask cache, M4 = MLP rows) stays in device memory and is read by the
consuming kernels; threshold compactions are a keep-flag and one index list.
Results are the same numbers as `voxurf_fine.Voxurf.forward` + autograd (tests/test_gpu_fused.py, test_gpu_fullsize.py).

Sequence (C entry point -> reference lines):
  vx_ray_setup                 t_min/t_max, N_steps, start/dir, scan           render_utils_kernel.cu:12-79,210-212
  vx_march_flags_cells / emit  bbox + MaskCache keep bits -> (ray_id, step_id) voxurf_fine.py:593-617,631-636,930-942
  vx_fused_sdf_alpha           sdf, 6-tap gradient, NeuS alpha, alpha > thres   voxurf_fine.py:640-648
  vx_alpha2weight_seg          T / weights (bit-exact recurrence), w > thres    voxurf_fine.py:667-669
  vx_scan_i32, vx_fused_emit_rows   row list                                    voxurf_fine.py:670-676
  vx_fused_row_features        k0 gather, sample_sdfs (L=4), PEs -> X1, X2      voxurf_fine.py:678-739
  vx_mlp_prep_batch            hi / lo TF32 weight images of both networks      (mlp.py, csrc/mlp_tc.cu)
  vx_mlp_chain_batch           rgbnet + k_rgbnet forward, one launch (tcgen05)  voxurf_fine.py:718,749
  vx_fused_composite_loss      sigmoid, segment sums, losses, their backward    voxurf_fine.py:752-763, run.py:604-636
  vx_mlp_chain_batch           the dX chains of both networks, one launch
  vx_mlp_dw_batch              all 8 weight / bias gradient GEMMs, one launch (side stream)
  vx_fused_row_backward        k0 scatter, sample_sdfs scatter                  (ATen grid_sampler backward)
  vx_alpha2weight_seg_backward                                                  render_utils_kernel.cu:653-677
  vx_fused_alpha_sdf_backward  NeuS alpha backward + 7-tap sdf scatter
  [TV iters] vx_fd_gradient_active, vx_smooth_grad_tv_masked_writes (side stream, beside the forward pass),
             vx_sdf_regularisers_backward(_slab)                                run.py:612-655
  vx_adam_step_blocklive (sdf), vx_adam_step_worklist (k0: voxels in touched | live), vx_adam_step x2 (MLPs)   lib/utils.py:154-199
defer_optimizer: the regulariser + Adam phase of step k runs at the start of step k + 1 beside its ray set-up and march (the k0
pass already inside step k, beside the rest of its backward pass).  world > 1: rays shard, optimizer state shards by X-slab, the
gradient / parameter exchange goes over NVLink peer memory (vx_block_nonzero, vx_pull_reduce, vx_k0_rows_scatter,
vx_adam_step_*_peers) with an NCCL fallback -- DESIGN.md section 7.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from ._lib import call
from .mlp import FlatMLP, prepare_chains, run_chain_jobs, run_dw_batch
from .optim import _storage


class FusedFineStep:
    def __init__(self, model, n_rays, train_cfg=None, render_kwargs=None, row_capacity=65536, world=1, rank=0,
                 tensor_core=True, sparse_k0_exchange=True, sparse_adam=True, use_graph=False,
                 graph_multi_gpu=True, dense_exchange=False, defer_optimizer=False, deterministic=False, k0_ownership=True):
        if model.k0_dim not in (6, 12):
            raise NotImplementedError('fused step: k0 channels must be 6 or 12')
        if model.k_center_sdf or not model.center_sdf or not model.k_res:
            raise NotImplementedError('fused step covers the shipped fine configs (center_sdf, k_res, no k_center_sdf)')
        self.m, self.N = model, int(n_rays)
        model._fused = self       # (Voxurf.mesh_color_forward reuses this step's flat MLPs and row buffers)
        self.cfg = dict(train_cfg) if train_cfg is not None else None
        self.rk = dict(render_kwargs or {})
        self.world, self.rank = world, rank
        self.tensor_core = tensor_core
        self.sparse_k0_exchange = sparse_k0_exchange
        dev = model.sdf.grid.device
        self.dev = dev
        m = model
        self.X, self.Y, self.Z = (int(w) for w in m.world_size)
        self.C = m.k0_dim
        self.k0_cl = int(m.k0.channels_last)
        self.disp = sorted(set(m.grad_feat + m.k_grad_feat))
        assert self.disp == sorted(set(m.sdf_feat + m.k_sdf_feat)) or len(m.k_sdf_feat) == 0
        self.disp = sorted(set(m.grad_feat))
        self.L = len(self.disp)
        self.P, self.Vp = m.posfreq.numel(), m.viewfreq.numel()
        self.P2, self.V2 = m.k_posfreq.numel(), m.k_viewfreq.numel()
        self.D1 = 3 + 6 * self.P + 3 + 6 * self.Vp + 1 + 9 * self.L
        self.D2 = self.C + 3 + 6 * self.P2 + 3 + 6 * self.V2 + 3 + 3
        self.ld1 = (self.D1 + 15) // 16 * 16
        self.ld2 = (self.D2 + 15) // 16 * 16
        self.col_logit = self.D2 - 3
        stepdist = float(np.float32(float(self.rk.get('stepsize', 0.5)) * m._voxel_size_host))
        self.stepdist = stepdist
        self.dist = float(np.float32(float(self.rk.get('stepsize', 0.5)) * m._voxel_size_host))
        N = self.N
        cap2 = N * m._max_steps(stepdist)
        self.cap2 = cap2
        f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        i32 = lambda *s: torch.empty(*s, dtype=torch.int32, device=dev)
        u8 = lambda *s: torch.empty(*s, dtype=torch.uint8, device=dev)
        # per ray
        self.t_min, self.t_max = f32(N), f32(N)
        self.n_steps = torch.empty(N, dtype=torch.int64, device=dev)
        self.start, self.dirs = f32(N, 3), f32(N, 3)
        self.offsets = torch.empty(N + 1, dtype=torch.int64, device=dev)
        words = N * (m._max_steps(stepdist) // 32 + 2) + 1
        self.bits_in, self.bits_keep = i32(words), i32(words)
        self.keep_count, self.keep_off = i32(N), i32(N + 1)
        self.w_count, self.off4 = i32(N), i32(N + 1)
        self.alphainv_last, self.i_end, self.d_last, self.loss_ray = f32(N), i32(N), f32(N), f32(N)
        self.rgb_marched, self.rgb_marched0 = f32(N, 3), f32(N, 3)
        self.normal_marched, self.depth = f32(N, 3), f32(N)
        self.loss = f32(1)
        # per M2 sample
        self.ray_id, self.step_id = i32(cap2), i32(cap2)
        self.sdf_s, self.grad_s, self.alpha = f32(cap2), f32(cap2, 3), f32(cap2)
        self.keep, self.w_keep = u8(cap2), u8(cap2)
        self.weight, self.T = f32(cap2), f32(cap2)
        self.d_w, self.d_alpha, self.d_sdf_s, self.d_grad_s = f32(cap2), f32(cap2), f32(cap2), f32(cap2, 3)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        # row-capacity overflow is polled without a sync: every `overflow_check_every` steps the counter is copied to
        # pinned memory behind an event, and the copy is inspected once the event has completed (a step or two later)
        self.overflow_check_every = 16
        self._ovf_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._ovf_event, self._ovf_pending, self._ovf_calls = torch.cuda.Event(), False, 0
        self._alloc_rows(int(row_capacity))
        # MLPs on flat parameter / gradient storage (one Adam launch per network)
        self.mlp1 = FlatMLP(m.rgbnet, self.ld1, self.D1, tensor_core)
        self.mlp2 = FlatMLP(m.k_rgbnet, self.ld2, self.D2, tensor_core)
        self.mlp1.alloc(self.cap4), self.mlp2.alloc(self.cap4)
        # one gradient buffer for both networks: the data-parallel exchange is one all-reduce for both
        n1, n2 = self.mlp1.flat.numel(), self.mlp2.flat.numel()
        self.mlp_grads = torch.zeros(n1 + n2, dtype=torch.float32, device=dev)
        self.mlp1.rebind_grad(self.mlp_grads[:n1]); self.mlp2.rebind_grad(self.mlp_grads[n1:])
        # persistent gradient buffers for the grids (zeroed by the Adam kernel itself)
        self.sdf_grad = torch.zeros_like(m.sdf.grid)
        self.k0_grad = torch.zeros_like(m.k0.grid, memory_format=torch.preserve_format)
        m.sdf.grid.grad, m.k0.grid.grad = self.sdf_grad, self.k0_grad
        # k0 bitmaps, one bit per voxel (channels-last: a voxel's C channels are contiguous).  touched: a gradient was
        # scattered there this step; live: one ever was.  The k0 Adam pass skips the gradient read where untouched and
        # the whole voxel where not live (m = v = g = 0: the dense update is the identity) -- same result, bit for bit,
        # as utils.Adam over the dense grid (lib/utils.py:154-199), far less traffic.  Only the scatter kernels write
        # k0 gradients in the fine stage (weight_tv_k0 = 0, configs/default_fine_s.py).
        self.k0_touched = self.k0_live = self._k0_list = None
        if self.k0_cl and sparse_adam:
            n_words = (self.X * self.Y * self.Z + 31) // 32
            self.k0_touched = torch.zeros(n_words, dtype=torch.int32, device=dev)
            self.k0_live = torch.zeros(n_words, dtype=torch.int32, device=dev)
            self._k0_list = torch.zeros(self.X * self.Y * self.Z + 1, dtype=torch.int32, device=dev)      # vx_adam_step_worklist scratch
        # deterministic=True: every scatter of the backward pass (k0 rows, sdf taps, split-K weight gradients) accumulates in
        # 64-bit fixed point (order-independent integer sums) and is folded into the fp32 gradient buffers afterwards:
        # gradients -- and with them the whole training trajectory -- are bit-reproducible run to run.  Costs one dense
        # read of the sdf accumulators per step (~25 us at 256^3) and 64-bit atomics.
        self.deterministic = bool(deterministic)
        self.ACC_SCALE, self.ACC_SCALE_MLP = float(2 ** 52), float(2 ** 44)
        if self.deterministic:
            assert world == 1, 'deterministic mode: single GPU (the NCCL reductions have their own order)'
            self.sdf_acc = torch.zeros(m.sdf.grid.numel(), dtype=torch.int64, device=dev)
            self.k0_acc = torch.zeros(m.k0.grid.numel(), dtype=torch.int64, device=dev)
            self.mlp_acc = torch.zeros(self.mlp_grads.numel(), dtype=torch.int64, device=dev)
        self.conv_scratch = None
        self._dw_stream = None
        self.G = None       # FD gradient grid + its gradient, allocated on the first TV iteration
        self.smoothed = self.d_smoothed = None
        if m.smooth_sdf:
            self.smoothed, self.d_smoothed = torch.empty_like(m.sdf.grid), torch.zeros_like(m.sdf.grid)
        self.adam_state = {}
        self.adam_steps = 0
        # CUDA-graph replay of the whole step (single GPU): static input buffers, the step-dependent scalars (1/s of the
        # NeuS schedule, Adam step sizes / bias corrections) in a small device array the kernels read (inv_s_dev, step_dev)
        # (world > 1: the NCCL collectives of the step are captured with it -- opt-in, graph_multi_gpu=True)
        self.use_graph = bool(use_graph) and (world == 1 or bool(graph_multi_gpu))
        self._graphs, self._eager_seen, self._dev_consts = {}, set(), None
        self._graph_launches, self.launches_replayed = {}, 0   # kernels per captured variant / total replayed (bench.py)
        if self.use_graph:
            self.consts = torch.zeros(16, dtype=torch.float32, device=dev)
            self.in_o, self.in_d, self.in_v, self.in_t = (torch.zeros(n_rays, 3, dtype=torch.float32, device=dev) for _ in range(4))
        self.force_eager = False   # bench.py: launch the kernels of the step one by one even when use_graph is set
        # defer_optimizer: the regulariser application, the gradient-exchange completion and the Adam passes of step k run at
        # the START of step k + 1, on a side stream beside that step's ray set-up and march (which read only the rays and the
        # mask cache, and are issue-bound while the optimizer is bandwidth / atomics-bound), and are joined before the first
        # read of the sdf grid.  Same arithmetic in the same order; parameters lag one call behind until the next step() or
        # flush() / sync_params().
        self.defer_optimizer = bool(defer_optimizer)
        self._pending = None
        self._opt_stream = None
        self._k0_work, self._works = [], []      # outstanding NCCL work of the data-parallel exchange
        self._k0_stream = None
        self._k0_hook, self._k0_forked = None, False
        # Slab-sharded data-parallel exchange (SURVEY.md 8e "preferred form"): reduce-scatter of the sdf gradient over X-slabs,
        # regularisers + Adam on the owned slab only, all-gather of the updated sdf PARAMETERS at the start of the next step
        # (it overlaps ray set-up and the march, which read rays and the mask cache only).  Non-owned slabs of this rank's sdf
        # grid / moments are stale between a step and the next all-gather: sync_params() / unshard() bring them up to date.
        self.sharded = world > 1 and not dense_exchange and self.X % world == 0 and not m.smooth_sdf
        self._params_dirty = False
        self._ag_work = None
        self._opt_forked = False
        if self.sharded:
            per = self.X // world * self.Y * self.Z
            self.slab_x = (rank * (self.X // world), (rank + 1) * (self.X // world))
            self.slab = (rank * per, (rank + 1) * per)
        # sdf Adam with block-level skipping (blocks of 128 voxels that never received a gradient are the identity)
        self.sdf_live = None
        if sparse_adam and m.sdf.grid.numel() % 128 == 0 and (not self.sharded or (self.slab[1] - self.slab[0]) % 128 == 0):
            self.sdf_live = torch.zeros(m.sdf.grid.numel() // 128, dtype=torch.uint8, device=dev)
        # k0 slab ownership (sharded step, channels-last k0 with bitmaps): each rank scatters only the rows' corners that fall
        # into its X-slab, steps only those voxels, and stores the updated parameters straight into the other ranks' replicas
        # of the k0 grid over NVLink peer memory (vx_adam_step_worklist_peers) -- re-scatter atomics and Adam traffic drop
        # by the world size and the replicas become bit-identical.  Needs CUDA IPC + peer access between the ranks' GPUs.
        # The sdf grid goes the same way (sdf_peer): instead of NCCL's dense reduce-scatter + all-gather (2 x 56 MB per rank and
        # step at 8 GPUs, the term that bounded the 8-GPU step), the owner of a slab pulls only the non-zero 128-voxel blocks
        # of every rank's gradient (vx_block_nonzero / vx_pull_reduce) and its block-live Adam stores the updated blocks into
        # the peers' replicas (vx_adam_step_blocklive_peers).
        self.k0_owned = self.sdf_peer = False
        self.k0_peer_note = None
        if self.sharded and k0_ownership and self.k0_touched is not None and sparse_k0_exchange and self.C in (6, 12):
            self._setup_peers()
        self.bitmap_probe = None   # bench.py: list collecting copies of (touched, live) as the k0 Adam launch sees them
        self.timings = None   # bench.py: list collecting (group, (start, end) CUDA events) around the k0 / sdf Adam launches
        if self.cfg is not None:
            c = self.cfg
            self.groups = [('sdf', [m.sdf.grid], c['lrate_sdf']), ('k0', [m.k0.grid], c['lrate_k0']),
                           ('rgbnet', [self.mlp1.flat], c['lrate_rgbnet']), ('k_rgbnet', [self.mlp2.flat], c['lrate_k_rgbnet'])]
            self.lr = {name: lr for name, _, lr in self.groups}

    def _setup_peers(self):
        """Map the other ranks' replicas of the k0 parameters, the sdf parameters, the sdf gradient and its block mask into this
        process (CUDA IPC handles of the containing cudaMalloc segments, exchanged over the process group and opened with this
        rank's device current: vx_ipc_*), so that this rank's kernels can store to / load from them over NVLink.  All ranks or
        none: on any failure the step keeps the NCCL exchange and the replicated k0 update."""
        import ctypes
        import torch.distributed as dist
        from ._lib import library, last_error
        W, r = self.world, self.rank
        lib = library()
        m = self.m
        self.sdf_mask = torch.zeros(m.sdf.grid.numel() // 128, dtype=torch.uint8, device=self.dev)
        tensors = {'k0': _storage(m.k0.grid.data).view(-1), 'sdf': m.sdf.grid.data.view(-1), 'sdf_grad': self.sdf_grad.view(-1),
                   'sdf_mask': self.sdf_mask}
        ok, note, mine = 1, '', None
        try:
            per = self.slab[1] - self.slab[0]
            if per % 128 or m.sdf.grid.numel() % 128:
                raise RuntimeError('slab size is not a multiple of the 128-voxel blocks')
            cur = torch.cuda.current_device()
            segs = [(sg['address'], sg['total_size']) for sg in torch.cuda.memory_snapshot() if sg['device'] == cur]
            mine = {}
            for name, t in tensors.items():
                ptr = t.data_ptr()
                base = [a for a, n in segs if a <= ptr and ptr + t.numel() * t.element_size() <= a + n]
                if len(base) != 1:
                    raise RuntimeError(f'{name} is not inside one cudaMalloc segment of the caching allocator')
                buf = (ctypes.c_uint8 * 64)()
                if lib.vx_ipc_get_handle(ctypes.c_void_p(base[0]), ctypes.cast(buf, ctypes.c_void_p)) != 0:
                    raise RuntimeError(last_error())
                mine[name] = (bytes(buf), ptr - base[0], t.numel())
        except Exception as e:      # noqa: BLE001 -- capability probe: IPC / P2P may be unavailable (containers, MIG, PCIe boxes)
            ok, note, mine = 0, f'{type(e).__name__}: {e}', None
        handles = [None] * W
        dist.all_gather_object(handles, mine)
        ptrs = {name: [t.data_ptr()] * W for name, t in tensors.items()}      # own address at index `rank`
        self._ipc_opened = []
        if ok and all(h is not None for h in handles):
            try:
                for q in range(W):
                    if q == r:
                        continue
                    bases = {}
                    for name, (hb, off, n) in handles[q].items():
                        assert n == tensors[name].numel()
                        if hb not in bases:       # (two tensors may share a segment: a handle is opened once per process)
                            hbuf = (ctypes.c_uint8 * 64).from_buffer_copy(hb)
                            out = ctypes.c_uint64(0)
                            if lib.vx_ipc_open_handle(ctypes.cast(hbuf, ctypes.c_void_p), ctypes.cast(ctypes.pointer(out), ctypes.c_void_p)) != 0:
                                raise RuntimeError(last_error())
                            bases[hb] = int(out.value)
                            self._ipc_opened.append(bases[hb])
                        ptrs[name][q] = bases[hb] + off
            except Exception as e:      # noqa: BLE001
                ok, note = 0, f'{type(e).__name__}: {e}'
        else:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)       # all ranks or none
        if int(flag.item()) != 1:
            self._close_k0_peers()
            self.k0_peer_note = 'peer memory unavailable (%s): NCCL exchange, replicated k0 update' % (note or 'a peer failed')
            return
        lo, per, C = self.slab[0], self.slab[1] - self.slab[0], self.C
        others = [q for q in range(W) if q != r]
        self._k0_peer_ptrs = [ptrs['k0'][q] + 4 * lo * C for q in others]        # my slab inside the peers' k0 replicas
        self._sdf_peer_ptrs = [ptrs['sdf'][q] + 4 * lo for q in others]          # ... inside their sdf parameter replicas
        self._sdf_grad_ptrs = [ptrs['sdf_grad'][q] + 4 * lo for q in range(W)]   # my slab of every rank's gradient (own included)
        self._sdf_mask_ptrs = [ptrs['sdf_mask'][q] + lo // 128 for q in range(W)]
        self._bar = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self.k0_owned = True
        self.sdf_peer = bool(self.sdf_live is not None)      # (the peer-store Adam is the block-live kernel)

    def _barrier(self, async_op=False):
        """cross-rank barrier on the current stream (a 4-byte NCCL all-reduce): every rank's earlier kernels on its stream --
        peer loads and stores included -- are complete when it completes"""
        import torch.distributed as dist
        return dist.all_reduce(self._bar, async_op=async_op)

    def _close_k0_peers(self):
        from ._lib import library
        for p in getattr(self, '_ipc_opened', []):
            library().vx_ipc_close_handle(p)
        self._ipc_opened = []

    # the model keeps a back-reference to its step (model._fused): copies / pickles of the model do not drag the step's
    # buffers, events and CUDA graphs along
    def __deepcopy__(self, memo):
        return None

    def __reduce__(self):
        return (type(None), ())

    def _alloc_rows(self, cap):
        dev = self.dev
        self.cap4 = cap
        self.idx4 = torch.empty(cap, dtype=torch.int32, device=dev)
        self.X1 = torch.empty(cap, self.ld1, dtype=torch.float32, device=dev)
        self.X2 = torch.empty(cap, self.ld2, dtype=torch.float32, device=dev)
        self.dX1, self.dX2 = torch.empty_like(self.X1), torch.empty_like(self.X2)
        self.logit1 = torch.empty(cap, 3, dtype=torch.float32, device=dev)
        self.k_out = torch.empty(cap, 3, dtype=torch.float32, device=dev)
        self.d_logit1, self.d_kout = torch.zeros_like(self.logit1), torch.zeros_like(self.k_out)
        if hasattr(self, 'mlp1'):
            self.mlp1.alloc(cap), self.mlp2.alloc(cap)

    # ------------------------------------------------------------------ forward pieces
    def _geom(self):
        m = self.m
        return (self.X, self.Y, self.Z, m._min_host, m._max_host)

    def _pts(self):
        return (self.ray_id, self.step_id, self.start, self.dirs, self.stepdist)

    def _forward(self, rays_o, rays_d, viewdirs, global_step, train, target=None):
        m, N = self.m, self.N
        assert rays_o.shape[0] == N, 'FusedFineStep is sized for a fixed batch; build another for a different one'
        rays_o, rays_d, viewdirs = rays_o.contiguous(), rays_d.contiguous(), viewdirs.contiguous()
        if self._dev_consts is None:
            s_val, inv_s = m._update_s_val(global_step)
            inv_s_dev = None
        else:                      # graph capture / replay: the host bookkeeping happened in _step_graph
            s_val, inv_s, inv_s_dev = 0, 0.0, self._dev_consts[0:1]
        self.inv_s, self._inv_s_dev = inv_s, inv_s_dev
        X, Y, Z, mn, mx = self._geom()
        near = self.rk['near']
        self._begin_param_gather()      # (sharded data-parallel step) runs under ray set-up and the march
        call('vx_ray_setup', rays_o, rays_d, m.xyz_min, m.xyz_max, near, 1e9, self.stepdist, N, self.t_min, self.t_max,
             self.n_steps, self.start, self.dirs, self.offsets)
        mc = m.mask_cache.march_args_cells() if m.mask_cache is not None else (None, 1, 1, 1, [0., 0., 0.], [1., 1., 1.], 0., 1., 0., None)
        call('vx_march_flags_cells', self.start, self.dirs, m.xyz_min, m.xyz_max, self.offsets, N, self.stepdist, *mc,
             self.bits_in, self.bits_keep, self.keep_count, self.keep_off)
        call('vx_march_emit', self.offsets, N, self.bits_keep, self.keep_off, self.cap2, self.ray_id, self.step_id, None)
        n2 = self.keep_off[N:]
        self._join_optimizer()          # (defer_optimizer) the previous step's Adam passes ran beside the march
        self._finish_param_gather()     # first read of the sdf grid follows
        if m.smooth_sdf:
            if self.conv_scratch is None:
                self.conv_scratch = torch.empty(2 * m.sdf.grid.numel(), dtype=torch.float32, device=self.dev)
            call('vx_conv3d_replicate_separable', m.sdf.grid, 1, X, Y, Z, m.smooth_conv.weight1d_host, m.smooth_conv.ksize, 0, 0,
                 self.conv_scratch, self.smoothed)
            sdf_grid = self.smoothed
        else:
            sdf_grid = m.sdf.grid
        self._sdf_grid = sdf_grid
        thres = float(m.fast_color_thres)
        call('vx_fused_sdf_alpha', sdf_grid, X, Y, Z, mn, mx, *self._pts(), n2, viewdirs, m._voxel_size_host, self.dist,
             inv_s, thres, self.sdf_s, self.grad_s, self.alpha, self.keep, self.d_w, self.d_sdf_s, self.d_grad_s, inv_s_dev)
        call('vx_alpha2weight_seg', self.alpha, self.keep if thres > 0 else None, self.keep_off, N, thres, self.weight, self.T,
             self.alphainv_last, self.i_end, self.w_keep, self.w_count)
        call('vx_scan_i32', self.w_count, N, self.off4)
        call('vx_fused_emit_rows', self.w_keep, self.keep_off, self.off4, N, self.cap4, self.idx4, self.overflow)
        n4 = self.off4[N:]
        call('vx_fused_row_features', sdf_grid, _storage(m.k0.grid), X, Y, Z, self.C, self.k0_cl, mn, mx, *self._pts(),
             self.idx4, n4, self.cap4, viewdirs, self.sdf_s, self.grad_s, m._voxel_size_host, int(m.use_grad_norm),
             self.P, self.Vp, self.P2, self.V2, self.disp, self.L, self.ld1, self.ld2, self.X1, self.X2)
        if self.tensor_core:   # all weight images of both networks (forward + transposed dX chains) in one launch
            prepare_chains(self.mlp1.chains(train) + self.mlp2.chains(train))
        if self.tensor_core:
            # both forward chains in one launch: rgbnet's tiles first, k_rgbnet's tiles take their rgb_logit.detach() input
            # columns (lib/voxurf_fine.py:741-751) straight from rgbnet's output rows behind a per-tile flag
            self.mlp1._n = self.mlp2._n = n4
            run_chain_jobs([self.mlp1.forward_job(self.X1, self.logit1, train),
                            self.mlp2.forward_job(self.X2, self.k_out, train, patch=(self.logit1, self.col_logit, 3, 0))],
                           n4, self.cap4, self.mlp1.done)
        else:
            self.mlp1.forward(self.X1, self.logit1, keep_activations=train, n_rows_dev=None, prepared=True)
            call('vx_fused_fill_logit_cols', self.logit1, 3, n4, self.cap4, self.col_logit, self.ld2, self.X2)
            self.mlp2.forward(self.X2, self.k_out, keep_activations=train, n_rows_dev=None, prepared=True)
        return s_val, n2, n4

    def _loss_cfg(self):
        c = self.cfg or {}
        w_ent = c.get('weight_entropy_last', 0.0)
        # run.py:607-610 reads the LAST ray of the batch only; with rays sharded over ranks that is the last ray of
        # the last rank, and gradients are averaged over ranks afterwards, hence the factor `world` there, 0 elsewhere
        ent_scale = 1.0 if self.world == 1 else (float(self.world) if self.rank == self.world - 1 else 0.0)
        return c.get('weight_main', 1.0), c.get('weight_rgb0', 0.0), w_ent, ent_scale

    # ------------------------------------------------------------------ public API
    @torch.no_grad()
    def render(self, rays_o, rays_d, viewdirs, render_grad=True, render_depth=True):
        """Inference forward (run.py:123-126 calls model(...) per 8192-ray chunk): -> dict of (N, .) tensors.
        Buffers are reused by the next call: clone what you keep."""
        if not torch.cuda.is_current_stream_capturing():
            self.flush()
        s_val, n2, n4 = self._forward(rays_o, rays_d, viewdirs, None, train=False)
        if not torch.cuda.is_current_stream_capturing():
            self.poll_overflow()
        call('vx_fused_composite_loss', self.logit1, self.k_out, 3, self.idx4, self.off4, self.cap4, self.weight,
             self.alphainv_last, None, self.N, 0.0, 0.0, 0.0, 0.0, float(self.rk.get('bg', 0.0)), 0, self.rgb_marched,
             self.rgb_marched0, None, None, None, None, None)
        if render_grad or render_depth:
            call('vx_fused_composite_aux', self.idx4, self.off4, self.cap4, self.weight, self.grad_s, self.step_id, self.dist,
                 self.N, self.normal_marched if render_grad else None, self.depth if render_depth else None)
        return {'rgb_marched': self.rgb_marched, 'rgb_marched0': self.rgb_marched0, 'alphainv_cum': self.alphainv_last,
                'normal_marched': self.normal_marched if render_grad else None, 'depth': self.depth if render_depth else None,
                'disp': (1 / self.depth) if render_depth else 0, 's_val': s_val}

    @torch.no_grad()
    def mesh_colors(self, pts):
        """Vertex colours of an extracted mesh, lib/voxurf_fine.py:804-892 (`mesh_color_forward`; run.py:889-893 feeds it the
        mesh vertices in chunks): the fine forward's feature build and both colour MLPs at arbitrary points, viewed along the
        inward surface normal (viewdirs = -normal, normal = gradient / (|gradient| + 1e-5)), no compositing.  Runs on the
        same kernels as a training step: every point is a one-sample "ray" (start = the point, zero direction), so
        vx_sdf_taps / vx_fused_row_features / the tcgen05 chains are used unchanged.  pts (P,3) -> rgb (P,3)"""
        from . import ops
        m = self.m
        self.sync_params()
        pts = pts.reshape(-1, 3).float().contiguous()
        P = pts.shape[0]
        out = torch.empty(P, 3, dtype=torch.float32, device=self.dev)
        X, Y, Z, mn, mx = self._geom()
        sdf_grid = m.smooth_conv(m.sdf.grid) if m.smooth_sdf else m.sdf.grid
        prepare_chains(self.mlp1.chains(False) + self.mlp2.chains(False))
        for a in range(0, P, self.cap4):
            chunk = pts[a:a + self.cap4].contiguous()
            n = chunk.shape[0]
            sdf, feat, grad = ops.sdf_taps(sdf_grid, chunk, mn, mx, [1.0], m._voxel_size_host, use_grad_norm=False, xyz_order=True,
                                           want_sdf=True)
            viewdirs = (-(grad / (grad.norm(dim=-1, keepdim=True) + 1e-5))).contiguous()
            ids = torch.arange(n, dtype=torch.int32, device=self.dev)
            zi, zd = torch.zeros(n, dtype=torch.int32, device=self.dev), torch.zeros(n, 3, dtype=torch.float32, device=self.dev)
            n_dev = torch.tensor([n], dtype=torch.int32, device=self.dev)
            call('vx_fused_row_features', sdf_grid, _storage(m.k0.grid), X, Y, Z, self.C, self.k0_cl, mn, mx, ids, zi, chunk, zd, 1.0,
                 ids, n_dev, self.cap4, viewdirs, sdf.contiguous(), grad.contiguous(), m._voxel_size_host, int(m.use_grad_norm),
                 self.P, self.Vp, self.P2, self.V2, self.disp, self.L, self.ld1, self.ld2, self.X1, self.X2)
            self.mlp1._n = self.mlp2._n = n_dev
            run_chain_jobs([self.mlp1.forward_job(self.X1, self.logit1, False),
                            self.mlp2.forward_job(self.X2, self.k_out, False, patch=(self.logit1, self.col_logit, 3, 0))],
                           n_dev, self.cap4, self.mlp1.done)
            out[a:a + n] = torch.sigmoid(self.logit1[:n] + self.k_out[:n])
        return out

    @torch.no_grad()
    def render_chunk(self, rays_o, rays_d, viewdirs, render_grad=True, render_depth=True):
        """render() of one chunk as ONE CUDA-graph replay (use_graph=True; run.py:123-126 renders a view as ~79 such
        chunks): static input buffers, results in the persistent output buffers (clone what you keep).  The NeuS
        sharpness is a launch constant of the captured kernels: the graph is re-captured if s_val changed."""
        if not self.use_graph or self.world > 1 and self.sharded and self._params_dirty:
            return self.render(rays_o, rays_d, viewdirs, render_grad, render_depth)
        key = (bool(render_grad), bool(render_depth), float(self.m._s_val_host) if hasattr(self.m, '_s_val_host') else None)
        rg = getattr(self, '_render_graph', None)
        if rg is None or rg[0] != key:
            self.render(rays_o, rays_d, viewdirs, render_grad, render_depth)     # eager once: lazy allocations, kernel attributes
            torch._foreach_copy_([self.in_o, self.in_d, self.in_v], [rays_o, rays_d, viewdirs])
            from ._lib import launch_count
            g = torch.cuda.CUDAGraph()
            l0 = launch_count()
            with torch.cuda.graph(g):
                out = self.render(self.in_o, self.in_d, self.in_v, render_grad, render_depth)
            self._render_graph = rg = (key, g, out, launch_count() - l0)
        torch._foreach_copy_([self.in_o, self.in_d, self.in_v], [rays_o, rays_d, viewdirs])
        rg[1].replay()
        self.launches_replayed += rg[3]
        self.poll_overflow()
        return rg[2]

    @torch.no_grad()
    def forward_backward(self, rays_o, rays_d, viewdirs, target, global_step):
        """Forward + losses + backward into the persistent gradient buffers.  Returns the (device) scalar loss."""
        m, N = self.m, self.N
        s_val, n2, n4 = self._forward(rays_o, rays_d, viewdirs, global_step, train=True)
        X, Y, Z, mn, mx = self._geom()
        w_main, w_rgb0, w_ent, ent_scale = self._loss_cfg()
        call('vx_fused_composite_loss', self.logit1, self.k_out, 3, self.idx4, self.off4, self.cap4, self.weight,
             self.alphainv_last, target.contiguous(), N, w_main, w_rgb0, w_ent, ent_scale, float(self.rk.get('bg', 0.0)), 1,
             self.rgb_marched, self.rgb_marched0, self.d_logit1, self.d_kout, self.d_w, self.d_last, self.loss_ray)
        call('vx_sum_f32', self.loss_ray, N, self.loss)
        sparse_dp = self.world > 1 and self.sparse_k0_exchange
        if self.tensor_core:   # both dX chains in one launch, then the 8 weight-gradient GEMMs of both networks in one launch
            run_chain_jobs([self.mlp2.backward_job(self.d_kout, self.dX2), self.mlp1.backward_job(self.d_logit1, self.dX1)],
                           n4, self.cap4, None)
            if sparse_dp:      # dL/dk0 of this rank's rows is final here: its all-gather runs under the rest of the backward
                self._start_k0_exchange(n4)
                if self._k0_hook is not None:
                    self._k0_hook()      # (deferred step) ... and so do the row scatter and the k0 Adam pass, on their own stream
            # The weight-gradient launch (one persistent CTA per SM, tensor / latency bound, ~220 us) only feeds the
            # optimizer; the scatter kernels below (atomics / ALU bound, independent inputs) run beside it on the main
            # stream.  Fork here, join at the end of this method; inside a CUDA-graph capture the side stream joins
            # the capture through the same wait_stream calls.
            if self._dw_stream is None:
                self._dw_stream = torch.cuda.Stream(device=self.dev)
            main = torch.cuda.current_stream()
            self._dw_stream.wait_stream(main)
            with torch.cuda.stream(self._dw_stream):
                if self.deterministic:
                    run_dw_batch([self.mlp2, self.mlp1], fx=(self.mlp_acc, self.mlp_grads, self.ACC_SCALE_MLP))
                    call('vx_fx_accumulate', self.mlp_acc, self.mlp_acc.numel(), self.ACC_SCALE_MLP, self.mlp_grads, None, 1)
                else:
                    run_dw_batch([self.mlp2, self.mlp1])
        else:
            self.mlp2.backward(self.d_kout, self.dX2)
            if sparse_dp:
                self._start_k0_exchange(n4)
            self.mlp1.backward(self.d_logit1, self.dX1)
        grad_target = self.d_smoothed if m.smooth_sdf else self.sdf_grad
        thres = float(m.fast_color_thres)
        if self.deterministic:
            call('vx_fused_row_backward_fx', self._sdf_grid, X, Y, Z, self.C, self.k0_cl, mn, mx, *self._pts(), self.idx4, n4,
                 self.cap4, m._voxel_size_host, int(m.use_grad_norm), self.P, self.Vp, self.P2, self.V2, self.disp, self.L,
                 self.ld1, self.ld2, self.dX1, self.dX2, self.d_sdf_s, self.d_grad_s, self.sdf_acc, self.k0_acc, self.ACC_SCALE,
                 self.k0_touched)
            call('vx_alpha2weight_seg_backward', self.alpha, self.weight, self.T, self.keep if thres > 0 else None,
                 self.alphainv_last, self.keep_off, self.i_end, N, self.d_w, self.d_last, self.d_alpha)
            call('vx_fused_alpha_sdf_backward_fx', X, Y, Z, mn, mx, *self._pts(), n2, viewdirs.contiguous(), self.sdf_s, self.grad_s,
                 self.keep, self.d_alpha, self.d_sdf_s, self.d_grad_s, m._voxel_size_host, self.dist, self.inv_s, self.sdf_acc,
                 self.ACC_SCALE, self._inv_s_dev)
            call('vx_fx_accumulate', self.sdf_acc, self.sdf_acc.numel(), self.ACC_SCALE, grad_target.view(-1), None, 1)
            call('vx_fx_accumulate', self.k0_acc, self.k0_acc.numel(), self.ACC_SCALE, _storage(self.k0_grad).reshape(-1),
                 self.k0_touched if self.k0_cl else None, self.C if self.k0_cl else 1)
        else:
            call('vx_fused_row_backward', self._sdf_grid, X, Y, Z, self.C, self.k0_cl, mn, mx, *self._pts(), self.idx4, n4,
                 self.cap4, m._voxel_size_host, int(m.use_grad_norm), self.P, self.Vp, self.P2, self.V2, self.disp, self.L,
                 self.ld1, self.ld2, self.dX1, self.dX2, self.d_sdf_s, self.d_grad_s, grad_target,
                 None if sparse_dp else _storage(self.k0_grad), None if sparse_dp else self.k0_touched)
            if self._k0_hook is not None and not sparse_dp:
                self._k0_hook()          # (deferred step, one GPU) the k0 gradient is final: its Adam pass runs beside the rest of the backward
            call('vx_alpha2weight_seg_backward', self.alpha, self.weight, self.T, self.keep if thres > 0 else None,
                 self.alphainv_last, self.keep_off, self.i_end, N, self.d_w, self.d_last, self.d_alpha)
            call('vx_fused_alpha_sdf_backward', X, Y, Z, mn, mx, *self._pts(), n2, viewdirs.contiguous(), self.sdf_s, self.grad_s,
                 self.keep, self.d_alpha, self.d_sdf_s, self.d_grad_s, m._voxel_size_host, self.dist, self.inv_s, grad_target,
                 self._inv_s_dev)
        if m.smooth_sdf:
            call('vx_conv3d_replicate_separable', self.d_smoothed, 1, X, Y, Z, m.smooth_conv.weight1d_host, m.smooth_conv.ksize, 1, 1,
                 self.conv_scratch, self.sdf_grad)
            self.d_smoothed.zero_()
        if self.tensor_core:
            torch.cuda.current_stream().wait_stream(self._dw_stream)   # join: the optimizer / exchange needs the weight gradients
        return self.loss

    # ------------------------------------------------------------------ data-parallel exchange
    def _start_k0_exchange(self, n4):
        """k0 gradient as rows: export (xyz, dk0 / world) of this rank's MLP rows and all-gather them asynchronously;
        the scatter of every rank's rows happens in grad_sync()."""
        import torch.distributed as dist
        W, cap, C = self.world, self.cap4, self.C
        if getattr(self, '_k0_send', None) is None or self._k0_send.numel() != cap * (3 + C) + 4:
            # one buffer per rank: [cap x 3 positions | cap x C gradient rows | row count] -> ONE all-gather
            self._k0_send = torch.empty(cap * (3 + C) + 4, dtype=torch.float32, device=self.dev)
            self._k0_recv = torch.empty(W, cap * (3 + C) + 4, dtype=torch.float32, device=self.dev)
        call('vx_fused_export_k0_rows', *self._pts(), self.idx4, n4, cap, self.dX2, self.ld2, C, 1.0 / W,
             self._k0_send[:cap * 3].view(cap, 3), self._k0_send[cap * 3:cap * (3 + C)].view(cap, C),
             self._k0_send[cap * (3 + C):].view(torch.int32))
        self._k0_work = [dist.all_gather_into_tensor(self._k0_recv.view(-1), self._k0_send, async_op=True)]

    def _begin_param_gather(self):
        """Sharded step: all-gather (in place) of the sdf slabs the ranks updated in the previous step."""
        if self.sharded and self._params_dirty and self._ag_work is None:
            import torch.distributed as dist
            flat = self.m.sdf.grid.data.view(-1)
            self._ag_work = dist.all_gather_into_tensor(flat, flat[self.slab[0]:self.slab[1]], async_op=True)

    def _finish_param_gather(self):
        if self._ag_work is not None:
            self._ag_work.wait()
            self._ag_work, self._params_dirty = None, False

    def gather_k0_grad(self):
        """k0 ownership: every rank holds the k0 gradient of its own X-slab only; all-gather the slabs (tests, parity_check)."""
        if self.k0_owned:
            import torch.distributed as dist
            flat = _storage(self.k0_grad).view(-1)
            lo, hi = self.slab[0] * self.C, self.slab[1] * self.C
            dist.all_gather_into_tensor(flat, flat[lo:hi].clone())

    def _gather_k0_state(self):
        """k0 ownership: Adam moments and live bits of a voxel exist on its owner only; all-gather the slabs so that every rank
        holds the full optimizer state (checkpoints, leaving the sharded mode)."""
        if not self.k0_owned:
            return
        import torch.distributed as dist
        C = self.C
        lo, hi = self.slab
        for t in (self.adam_state.get(id(self.m.k0.grid)) or ()):
            flat = _storage(t).view(-1)
            dist.all_gather_into_tensor(flat, flat[lo * C:hi * C].clone())
        dist.all_gather_into_tensor(self.k0_live, self.k0_live[lo // 32:hi // 32].clone())

    def shutdown(self):
        """Release captured graphs (they hold NCCL work) and the peers' memory mappings: call before destroying the process group."""
        self.flush()
        self.release_graphs()
        if self.k0_owned:
            import torch.distributed as dist
            self._gather_k0_state()
            self.k0_owned = self.sdf_peer = False
            torch.cuda.synchronize()
            self._close_k0_peers()            # unmap the peers' replicas ...
            dist.barrier()                    # ... before any owner may free its own

    def sync_params(self):
        """Make this rank's copy of the sdf grid current (sharded step: the slabs other ranks own are stale between a
        step and the next step's all-gather; with defer_optimizer the last step's update is still pending).  Call before
        reading parameters: rendering, checkpoints, evaluation."""
        self.flush()
        self._begin_param_gather()
        self._finish_param_gather()

    def unshard(self):
        """Leave the slab-sharded mode: parameters and Adam moments of the sdf grid are all-gathered so that every rank
        holds the full optimizer state again (checkpoints; regulariser combinations the slab kernels do not cover)."""
        if not self.sharded:
            return
        import torch.distributed as dist
        self.sync_params()
        st = self.adam_state.get(id(self.m.sdf.grid))
        if st is not None:
            for t in st:
                flat = t.view(-1)
                dist.all_gather_into_tensor(flat, flat[self.slab[0]:self.slab[1]])
        self._gather_k0_state()
        self.k0_owned = self.sdf_peer = False
        self.sharded = False
        if self.sdf_live is not None:
            self.sdf_live.fill_(1)       # the gathered moments of the other ranks' slabs may be non-zero anywhere
        self.release_graphs()

    def _sync_begin(self):
        import torch.distributed as dist
        AVG = dist.ReduceOp.AVG
        if self.sdf_peer:
            # peer-memory exchange: flag the non-zero blocks of this rank's gradient; the barrier (joined in _sync_end, before the
            # owners pull) says that every rank's backward pass and flags are complete
            call('vx_block_nonzero', self.sdf_grad.view(-1), self.sdf_grad.numel(), self.sdf_mask)
            self._works = [self._barrier(async_op=True)]
        elif self.sharded:
            flat = self.sdf_grad.view(-1)      # in place: the owned slab of this buffer receives the rank-averaged gradient
            self._works = [dist.reduce_scatter_tensor(flat[self.slab[0]:self.slab[1]], flat, op=AVG, async_op=True)]
        else:
            self._works = [dist.all_reduce(self.sdf_grad, op=AVG, async_op=True)]
        self._works += [dist.all_reduce(self.mlp_grads, op=AVG, async_op=True)]

    def _sync_k0(self):
        import torch.distributed as dist
        m = self.m
        if self.sparse_k0_exchange:
            for w in self._k0_work:
                w.wait()
            self._k0_work = []
            cap, C = self.cap4, self.C
            if self.k0_cl and C in (6, 12):
                # every rank's rows, as many as it sent, in one launch; with k0 ownership only the corners inside the owned slab
                x_lo, x_hi = self.slab_x if self.k0_owned else (0, self.X)
                call('vx_k0_rows_scatter', self.X, self.Y, self.Z, C, m._min_host, m._max_host, self._k0_recv.view(-1), self.world, cap,
                     x_lo, x_hi, _storage(self.k0_grad).view(-1), self.k0_touched)
                return
            for r in range(self.world):
                xyz, g = self._k0_recv[r, :cap * 3].view(cap, 3), self._k0_recv[r, cap * 3:cap * (3 + C)].view(cap, C)
                call('vx_grid_gather_backward', self.X, self.Y, self.Z, self.C, self.k0_cl, m._min_host, m._max_host, xyz, None, None,
                     None, None, 0.0, self._k0_recv[r, cap * (3 + C):].view(torch.int32), cap, g, _storage(self.k0_grad), self.k0_touched)
        else:
            dist.all_reduce(_storage(self.k0_grad), op=dist.ReduceOp.AVG)
            if self.k0_touched is not None:
                self.k0_touched.fill_(-1)   # dense exchange: every voxel may carry a gradient

    def _sync_end(self):
        for w in self._works:
            w.wait()
        self._works = []
        if self.sdf_peer:
            # this rank's slab <- mean over ranks of the flagged blocks, pulled over NVLink in rank order
            lo, hi = self.slab
            call('vx_pull_reduce', self.sdf_grad.view(-1)[lo:hi], hi - lo, self._sdf_grad_ptrs, self._sdf_mask_ptrs, self.world,
                 1.0 / self.world)

    def grad_sync(self):
        """Average gradients over ranks: dense NCCL all-reduce (AVG) for the sdf grid (67 MB at 256^3) and the two flat
        MLP buffers; the k0 grid (0.8 GB dense) is exchanged as ~3 MB of rows per rank and re-scattered locally."""
        if self.world <= 1:
            return
        self._sync_begin()
        self._sync_k0()
        self._sync_end()

    def is_tv_iter(self, global_step):
        c = self.cfg
        return c['tv_from'] < global_step < c['tv_end'] and global_step % c['tv_every'] == 0

    @torch.no_grad()
    def tv_flags(self, global_step):
        """(is a TV iteration with an active regulariser, dense TV add-grad) -- the two step-dependent branches"""
        c = self.cfg
        tv = self.is_tv_iter(global_step) and c['weight_tv_density'] > 0 and not c.get('ori_tv', False)
        return bool(tv), bool(global_step < c['tv_dense_before'])

    def regularise_prepare(self, flags):
        """First half of the smooth-gradient TV regulariser (run.py:612-625): FD gradient grid -> dL/dG and the loss
        term.  It only reads the sdf grid, so _step_body runs it on the side stream beside the forward / backward pass."""
        c, m = self.cfg, self.m
        is_tv, dense = flags
        tv = c['tv_terms']
        if not is_tv or tv['smooth_grad_tv'] <= 0:
            return
        X, Y, Z = self.X, self.Y, self.Z
        if self.G is None:
            # Only the part of the gradient grid within one voxel of the non-empty mask is ever read by the
            # regulariser (3^3 smoothing of masked voxels), and dL/dG is zero outside the mask: G and dG are
            # zero-initialised once, the FD gradient is evaluated inside the dilated mask only, dG is written inside
            # the mask only, and the FD adjoint is skipped outside the dilated mask.  Same numbers, ~85 % less
            # traffic for a typical mask.  (m.gradient is therefore only valid near the mask in this path.)
            self.G = torch.zeros(1, 3, X, Y, Z, dtype=torch.float32, device=self.dev)
            self.dG = torch.zeros_like(self.G)
            self.tv_active = (F.max_pool3d(m.nonempty_mask.float(), 3, 1, 1) > 0)[0, 0].contiguous()
            self.tv_loss = torch.zeros(1, dtype=torch.float32, device=self.dev)
            self.tv_scratch = torch.empty(int(call('vx_smooth_grad_tv_scratch_floats')), dtype=torch.float32, device=self.dev)
        call('vx_fd_gradient_active', m.sdf.grid, X, Y, Z, m._voxel_size_host, self.tv_active, self.G)
        m.gradient = self.G
        w = c['weight_tv_density'] * tv['smooth_grad_tv'] / (3.0 * m._n_nonempty)
        call('vx_smooth_grad_tv_masked_writes', self.G, m.nonempty_mask[0, 0], X, Y, Z, m._tv_smooth_w, w, self.dG, self.tv_scratch, self.tv_loss)

    def regularise_apply(self, flags, global_batch=None, add_loss=True):
        """Second half: the regularisers' gradients land in the sdf gradient (run.py:622-625, 641-655)."""
        c, m = self.cfg, self.m
        is_tv, dense = flags
        if not is_tv:
            return
        X, Y, Z = self.X, self.Y, self.Z
        tv = c['tv_terms']
        if tv['smooth_grad_tv'] > 0 and add_loss:
            self.loss.add_(self.tv_loss)   # run.py:622-625 adds the regulariser to the reported loss
        n_batch = global_batch or self.N * self.world
        wt = c['weight_tv_density'] * tv['sdf_tv'] / n_batch * max(X, Y, Z) / 128
        if self.sharded:    # only the owned X-slab of the gradient is regularised (and stepped); _slab_covers(flags) holds
            call('vx_sdf_regularisers_backward_slab', self.dG if tv['smooth_grad_tv'] > 0 else None, m.sdf.grid, X, Y, Z,
                 m._voxel_size_host, wt, wt, wt, self.sdf_grad, self.tv_active if tv['smooth_grad_tv'] > 0 else None, *self.slab_x)
            if self.sdf_peer:
                # the slab kernel read one halo plane of the neighbouring slabs from this rank's parameter replica: their owners
                # must not store updated parameters into it before every rank is past this point
                self._barrier()
            return
        if tv['smooth_grad_tv'] > 0 and tv['sdf_tv'] > 0 and dense:
            # both regularisers land in the sdf gradient with one read-modify-write
            call('vx_sdf_regularisers_backward', self.dG, m.sdf.grid, X, Y, Z, m._voxel_size_host, wt, wt, wt, self.sdf_grad,
                 self.tv_active)
            return
        if tv['smooth_grad_tv'] > 0:
            call('vx_fd_gradient_backward', self.dG, X, Y, Z, m._voxel_size_host, self.sdf_grad)
        if tv['sdf_tv'] > 0:
            call('vx_total_variation_add_grad', m.sdf.grid, self.sdf_grad, None, wt, wt, wt,
                 int(dense), X, Y, Z, m.sdf.grid.numel())

    def _slab_covers(self, flags):
        """Regulariser combinations the slab kernel implements: dense TV add-grad with or without the smooth-gradient TV."""
        is_tv, dense = flags
        if not is_tv:
            return True
        tv = self.cfg['tv_terms']
        return dense and tv['sdf_tv'] > 0 and self.X * self.Y * self.Z < 2 ** 32 and self.Z % 4 == 0

    def regularise(self, global_step, global_batch=None, flags=None):
        """run.py:612-625 (smooth-grad TV through the full-grid FD gradient) and run.py:641-655 (TV add-grad)."""
        flags = flags if flags is not None else self.tv_flags(global_step)
        self.regularise_prepare(flags)
        self.regularise_apply(flags, global_batch)

    @torch.no_grad()
    def optimizer_step(self, only=None, advance=True, step=None, lrs=None, barrier=True, early=False):
        """lib/utils.py:83-199 with betas (0.9, 0.99), eps 1e-8 (lib/utils.py:229); grads are zeroed in the same pass.
        only: restrict to these group names (the multi-GPU step updates k0 while the sdf all-reduce is in flight).
        step / lrs: the Adam step number and learning rates to apply (default: advance the counter, current rates)."""
        if step is None:
            if advance and self._dev_consts is None:
                self.adam_steps += 1
            step = max(self.adam_steps, 1)
        lrs = lrs if lrs is not None else self.lr
        beta1, beta2, eps = 0.9, 0.99, 1e-8
        bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
        peer_stores = False
        for gi, (name, params, _) in enumerate(self.groups):
            if only is not None and name not in only:
                continue
            lr = lrs[name]
            for p in params:
                st = self.adam_state.get(id(p))
                if st is None:
                    st = (torch.zeros_like(p, memory_format=torch.preserve_format), torch.zeros_like(p, memory_format=torch.preserve_format))
                    self.adam_state[id(p)] = st
                timed = self.timings is not None and name in ('k0', 'sdf')
                if timed:
                    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                    ev[0].record()
                touched, live = (self.k0_touched, self.k0_live) if name == 'k0' else (None, None)
                if touched is not None and self.bitmap_probe is not None:
                    self.bitmap_probe.append((touched.clone(), live.clone()))
                tensors = [_storage(p.data), _storage(p.grad), _storage(st[0]), _storage(st[1])]
                sdf_live = self.sdf_live if name == 'sdf' else None
                if name == 'sdf' and self.sharded:
                    # the owned X-slab only; the other slabs of the gradient buffer still hold this rank's local
                    # (un-reduced) gradient and are cleared here, the Adam pass clears the slab itself
                    lo, hi = self.slab
                    g = tensors[1].view(-1)
                    if not self.sdf_peer:
                        g[:lo].zero_(); g[hi:].zero_()
                    tensors = [t.view(-1)[lo:hi] for t in tensors]
                    sdf_live = sdf_live[lo // 128:hi // 128] if sdf_live is not None else None
                    self._params_dirty = not self.sdf_peer
                if sdf_live is not None and name == 'sdf' and self.sharded and self.sdf_peer:
                    # the updated blocks go straight into the peers' replicas; the other slabs of the gradient buffer (this rank's
                    # local gradient, pulled by their owners) are cleared behind the closing barrier below
                    call('vx_adam_step_blocklive_peers', *tensors, tensors[0].numel(), beta1, beta2, 1 - beta1, 1 - beta2, lr / bc1,
                         math.sqrt(bc2), eps, 1, sdf_live, None if self._dev_consts is None else self._dev_consts[2 + 2 * gi:4 + 2 * gi],
                         self._sdf_peer_ptrs, len(self._sdf_peer_ptrs))
                    peer_stores = 'sdf'
                elif sdf_live is not None:
                    call('vx_adam_step_blocklive', *tensors, tensors[0].numel(), beta1, beta2, 1 - beta1, 1 - beta2, lr / bc1,
                         math.sqrt(bc2), eps, 1, sdf_live, None if self._dev_consts is None else self._dev_consts[2 + 2 * gi:4 + 2 * gi])
                if sdf_live is not None:
                    if timed:
                        ev[1].record()
                        self.timings.append((name, ev))
                    continue
                if touched is not None and self.k0_owned:
                    # the owned X-slab only; the new parameters go to every replica (peer stores); cross-rank barrier below
                    lo, hi = self.slab
                    C = self.C
                    sl = [t.view(-1)[lo * C:hi * C] for t in tensors]
                    call('vx_adam_step_worklist_peers', *sl, sl[0].numel(), beta1, beta2, 1 - beta1, 1 - beta2, lr / bc1,
                         math.sqrt(bc2), eps, 1, touched[lo // 32:hi // 32], live[lo // 32:hi // 32], C, 1, self._k0_list,
                         None if self._dev_consts is None else (self._dev_consts[10:12] if early else self._dev_consts[2 + 2 * gi:4 + 2 * gi]),
                         self._k0_peer_ptrs, len(self._k0_peer_ptrs))
                    peer_stores = peer_stores or 'k0'
                elif touched is not None:
                    call('vx_adam_step_worklist', *tensors, tensors[0].numel(), beta1, beta2, 1 - beta1, 1 - beta2, lr / bc1,
                         math.sqrt(bc2), eps, 1, touched, live, self.C, 1, self._k0_list,
                         None if self._dev_consts is None else (self._dev_consts[10:12] if early else self._dev_consts[2 + 2 * gi:4 + 2 * gi]))
                else:
                    call('vx_adam_step', *tensors, None, tensors[0].numel(),
                         beta1, beta2, 1 - beta1, 1 - beta2, lr / bc1, math.sqrt(bc2), eps, 0, 1, touched, live, self.C,
                         None if self._dev_consts is None else self._dev_consts[2 + 2 * gi:4 + 2 * gi])
                    if touched is not None:
                        call('vx_bitmap_merge', live, touched, touched.numel())
                if timed:
                    ev[1].record()
                    self.timings.append((name, ev))
        if peer_stores and barrier:
            # every rank's peer stores have landed (and every owner has pulled the gradient blocks it reduces) before anyone
            # reads the parameters again / clears the gradient slabs it does not own
            self._barrier()
            if peer_stores == 'sdf':
                lo, hi = self.slab
                g = self.sdf_grad.view(-1)
                g[:lo].zero_(); g[hi:].zero_()

    def state_dict(self):
        """Optimizer state of the fused step (the model's own state_dict holds the parameters): Adam moments per
        group, step count, decayed learning rates, the k0 `live` bitmap.  Mirrors what run.py:786-793 saves as
        optimizer_state_dict."""
        self.sync_s_val()
        self.flush()
        if self.sharded:      # every rank saves the full sdf parameters and moments
            import torch.distributed as dist
            self.sync_params()
            mom = self.adam_state.get(id(self.m.sdf.grid))
            for t in (mom or ()):
                dist.all_gather_into_tensor(t.view(-1), t.view(-1)[self.slab[0]:self.slab[1]])
            self._gather_k0_state()
        st = {'adam_steps': self.adam_steps, 'lr': dict(self.lr), 'moments': {}}
        for name, params, _ in self.groups:
            m = self.adam_state.get(id(params[0]))
            if m is not None:
                st['moments'][name] = (m[0].detach().clone(), m[1].detach().clone())
        if self.k0_live is not None:
            st['k0_live'] = self.k0_live.clone()
        return st

    def load_state_dict(self, st):
        self.adam_steps = int(st['adam_steps'])
        self.lr.update(st['lr'])
        for name, params, _ in self.groups:
            if name in st['moments']:
                p = params[0]
                cur = self.adam_state.get(id(p))
                if cur is None:
                    cur = (torch.zeros_like(p, memory_format=torch.preserve_format), torch.zeros_like(p, memory_format=torch.preserve_format))
                    self.adam_state[id(p)] = cur
                for dst, src in zip(cur, st['moments'][name]):     # in place: captured graphs keep their pointers
                    _storage(dst).copy_(_storage(src.to(dst.device)) if src.shape == dst.shape else src.to(dst.device).reshape(_storage(dst).shape))
        if self.k0_live is not None:
            if 'k0_live' in st:
                self.k0_live.copy_(st['k0_live'])
            else:
                self.mark_all_live()    # moments of unknown sparsity: every voxel may hold non-zero exp_avg / exp_avg_sq
        if self.sdf_live is not None:
            self.sdf_live.fill_(1)
        self.m._refresh_derived()

    def warm_up(self, batches, global_step):
        """Run real training steps from `global_step` until every execution variant of the step (TV / non-TV iteration,
        and with defer_optimizer the variant of the preceding step) has had its eager first occurrence and -- with
        use_graph -- its CUDA graph captured, so that later steps are pure replays.  `batches`: list of
        (rays_o, rays_d, viewdirs, target).  Returns the next global_step."""
        n_look = 3 * max(1, self.cfg['tv_every']) + 2
        fl = [self.tv_flags(global_step + i) for i in range(1, n_look)]
        want = {(b, a if self.defer_optimizer else None) for a, b in zip(fl[:-1], fl[1:])}
        if not self.defer_optimizer:
            want = {(f, None) for f in fl}
        i = 0
        while not (set(self._graphs) >= want if self.use_graph else i >= 2) and i < 8 * max(1, self.cfg['tv_every']):
            self.step(*batches[i % len(batches)], global_step + i)
            i += 1
        return global_step + i

    def mark_all_live(self):
        """Call after loading optimizer moments from elsewhere: every voxel may then hold non-zero exp_avg / exp_avg_sq."""
        if self.k0_live is not None:
            self.k0_live.fill_(-1)
        if self.sdf_live is not None:
            self.sdf_live.fill_(1)

    def apply_lr_decay(self):
        """run.py:679-683: every learning rate times 0.1 ** (1 / (lrate_decay * 1000)) after each optimizer step."""
        f = 0.1 ** (1 / (self.cfg['lrate_decay'] * 1000))
        for k in self.lr:
            self.lr[k] *= f

    def _optimizer_phase(self, pend):
        """Everything between a step's backward pass and the next forward: completion of the gradient exchange, the k0
        row re-scatter, the regularisers' gradients, the Adam passes (run.py:641-659).  pend: dict(flags, step, lrs)."""
        flags, kw = pend['flags'], dict(step=pend['step'], lrs=pend['lrs'])
        k0_done = bool(pend.get('k0_done'))      # the k0 grid was stepped inside its own step's launch (_k0_early)
        if self.world > 1:
            if self.defer_optimizer:
                # the sdf exchange / MLP all-reduce of the pending step start HERE, not at the end of that step's
                # body: every collective then begins and ends inside one launch of the step (a CUDA-graph capture cannot
                # leave NCCL work unjoined, nor wait for work of an earlier launch), and they still run beside the next step's march
                self._sync_begin()
            if self.sdf_peer:
                # the k0 path (row scatter + owner-side Adam with peer stores) and the sdf path (barrier, pull-reduce, regulariser,
                # slab Adam with peer stores) are independent: side by side on two streams, one closing barrier for both grids
                cur = torch.cuda.current_stream()
                if not k0_done:
                    if self._k0_stream is None:
                        self._k0_stream = torch.cuda.Stream(device=self.dev)
                    self._k0_stream.wait_stream(cur)
                    with torch.cuda.stream(self._k0_stream):
                        self._sync_k0()
                        self.optimizer_step(only=('k0',), barrier=False, **kw)
                self._sync_end()
                self.regularise_apply(flags, add_loss=False)
                if not k0_done:
                    cur.wait_stream(self._k0_stream)
            else:
                if not k0_done:
                    self._sync_k0()
                    self.optimizer_step(only=('k0',), **kw)
                elif self.k0_owned:
                    self._barrier()      # the early k0 pass's peer stores
                self._sync_end()
                self.regularise_apply(flags, add_loss=False)
            self.optimizer_step(only=('sdf', 'rgbnet', 'k_rgbnet'), **kw)
        else:
            self.regularise_apply(flags, add_loss=False)
            self.optimizer_step(only=('sdf', 'rgbnet', 'k_rgbnet') if k0_done else None, **kw)

    def _k0_early_enabled(self):
        # One GPU only.  Measured at N > 1 (profiles/r02_bench_n8_k0_early_experiment.json): the k0 path then hangs off the row
        # all-gather, i.e. off the slowest rank's backward pass, and joining it at the end of the launch costs more than it hides
        # (N = 8: 1.26 -> 1.32 ms; N = 2: 1.156 -> 1.143 ms); there it stays in the deferred phase beside the next step's march.
        return bool(self.defer_optimizer and self.k0_touched is not None and self.k0_cl and not self.deterministic and self.tensor_core
                    and self.world == 1)

    def _k0_early(self, kw):
        """(deferred step) The k0 grid's share of the optimizer phase inside its own step's launch, as soon as its gradient is
        final -- after k_row_backward on one GPU; at N > 1 after the k0 rows are exported: all-gather, owned-corner scatter --
        then the voxel-list Adam pass (with the peer stores), on a stream of its own beside the rest of the backward pass, which
        neither reads nor writes the k0 grid.  The deferred phase of the next launch is then the sdf grid and the MLPs only."""
        main = torch.cuda.current_stream()
        if self._k0_stream is None:
            self._k0_stream = torch.cuda.Stream(device=self.dev)
        self._k0_stream.wait_stream(main)
        with torch.cuda.stream(self._k0_stream):
            if self.world > 1:
                self._sync_k0()
            self.optimizer_step(only=('k0',), barrier=False, early=True, **kw)
        self._k0_forked = True

    def flush(self):
        """Apply a deferred optimizer phase now (defer_optimizer=True): call before reading parameters."""
        if self._pending is not None:
            pend, self._pending = self._pending, None
            self._optimizer_phase(pend)

    def _join_optimizer(self):
        """forward pass, right before the first read of a parameter: the deferred optimizer phase must be complete"""
        if self._opt_forked:
            torch.cuda.current_stream().wait_stream(self._opt_stream)
            self._opt_forked = False

    def _step_body(self, rays_o, rays_d, viewdirs, target, global_step, flags, pend_in='own'):
        """pend_in: the optimizer phase to run at the start of this step ('own': whatever self._pending holds)."""
        if self.sharded and not self._slab_covers(flags):
            self.flush()
            self.unshard()      # e.g. sparse TV after tv_dense_before: back to the dense exchange, full optimizer state everywhere
        if pend_in == 'own':
            pend_in, self._pending = self._pending, None
        main = torch.cuda.current_stream()
        if self._dw_stream is None:
            self._dw_stream = torch.cuda.Stream(device=self.dev)
        if self._opt_stream is None:
            self._opt_stream = torch.cuda.Stream(device=self.dev)
        self._opt_forked = False
        if pend_in is not None:
            # the previous step's optimizer phase, beside this step's ray set-up and march
            self._opt_stream.wait_stream(main)
            with torch.cuda.stream(self._opt_stream):
                self._optimizer_phase(pend_in)
                self._begin_param_gather()       # (sharded) the all-gather of the updated slabs follows the Adam pass
            self._opt_forked = True
        early_tv = bool(flags[0] and self.tensor_core)
        if early_tv:
            # the first half of the TV regulariser depends on the parameters only: it runs on the side stream (ahead of
            # the weight-gradient launch that is forked onto the same stream later) beside the forward / backward pass
            self._begin_param_gather()
            self._dw_stream.wait_stream(main)
            if pend_in is not None:
                self._dw_stream.wait_stream(self._opt_stream)      # ... on the UPDATED parameters
            with torch.cuda.stream(self._dw_stream):
                if self._ag_work is not None:
                    self._ag_work.wait()     # the FD gradient reads the whole sdf grid
                self.regularise_prepare(flags)
        k0_early = self._k0_early_enabled()
        if k0_early:
            kw0 = dict(step=max(self.adam_steps + (1 if self._dev_consts is None else 0), 1), lrs=dict(self.lr))
            self._k0_hook = lambda: self._k0_early(kw0)
        try:
            loss = self.forward_backward(rays_o, rays_d, viewdirs, target, global_step)
        finally:
            self._k0_hook = None
        if self._k0_forked:
            main.wait_stream(self._k0_stream)      # join inside this launch
            self._k0_forked = False
        if not early_tv:
            self.regularise_prepare(flags)
        if flags[0] and self.cfg['tv_terms']['smooth_grad_tv'] > 0:
            self.loss.add_(self.tv_loss)   # run.py:622-625 adds the regulariser to the reported loss
        if self.world > 1:
            if self.defer_optimizer:
                for w in self._k0_work:      # the k0 row all-gather started in the backward pass: join it inside this launch
                    w.wait()
                self._k0_work = []
            else:
                self._sync_begin()   # the sdf reduce-scatter (or all-reduce) and the MLP all-reduce start right after the backward pass
        if self._dev_consts is None:
            self.adam_steps += 1
        pend = dict(flags=flags, step=max(self.adam_steps, 1), lrs=dict(self.lr), k0_done=k0_early)
        if self.defer_optimizer:
            self._pending = pend
        else:
            self._optimizer_phase(pend)
        return loss

    def _step_graph(self, rays_o, rays_d, viewdirs, target, global_step):
        """The same step as one CUDA-graph launch.  Per variant -- (TV iteration?, dense TV?) of this step and, with
        defer_optimizer, of the step whose optimizer phase runs inside this launch -- the first occurrence runs eagerly
        (lazy allocations, kernel attributes), the second is captured, later ones are replayed.  Host side per step: the
        NeuS schedule and the Adam scalars (a 64-byte upload) and one fused copy of the batch into the static input buffers."""
        m = self.m
        flags = self.tv_flags(global_step)
        pend_in = self._pending
        key = (flags, pend_in['flags'] if pend_in is not None else None)
        if key not in self._eager_seen:
            self._eager_seen.add(key)
            return self._step_body(rays_o, rays_d, viewdirs, target, global_step, flags)
        s_val = 1. / (global_step + m.s_ratio / m.s_start - m.step_start) * m.s_ratio          # lib/voxurf_fine.py:466-473
        m._s_val_host = float(np.float32(s_val))
        self.adam_steps += 1
        cur = dict(flags=flags, step=self.adam_steps, lrs=dict(self.lr), k0_done=self._k0_early_enabled())
        opt = pend_in if self.defer_optimizer else cur      # the optimizer phase that executes inside this launch
        host = [float(np.float32(1.0) / np.float32(m._s_val_host)), 0.0]
        for name, _, _ in self.groups:
            if opt is None:
                host += [0.0, 1.0]
            else:
                bc1, bc2 = 1 - 0.9 ** opt['step'], 1 - 0.99 ** opt['step']
                host += [opt['lrs'][name] / bc1, math.sqrt(bc2)]
        bc1, bc2 = 1 - 0.9 ** cur['step'], 1 - 0.99 ** cur['step']
        host += [cur['lrs']['k0'] / bc1, math.sqrt(bc2)]      # [10:12]: THIS step's k0 pass (_k0_early runs it inside this launch)
        self.consts[:len(host)].copy_(torch.tensor(host, dtype=torch.float32))
        torch._foreach_copy_([self.in_o, self.in_d, self.in_v, self.in_t], [rays_o, rays_d, viewdirs, target])
        g = self._graphs.get(key)
        if g is None:
            assert self.timings is None and self.bitmap_probe is None, 'kernel timing hooks need use_graph=False'
            self._dev_consts = self.consts
            g = torch.cuda.CUDAGraph()
            from ._lib import launch_count
            l0 = launch_count()
            try:
                with torch.cuda.graph(g):
                    self._step_body(self.in_o, self.in_d, self.in_v, self.in_t, None, flags, pend_in=pend_in)
            finally:
                self._dev_consts = None
            self._graphs[key] = g
            self._graph_launches[key] = launch_count() - l0
        g.replay()
        self._pending = cur if self.defer_optimizer else None
        self._params_dirty = self.sharded and not self.sdf_peer      # (host flags do not move during a replay)
        self.launches_replayed += self._graph_launches[key]
        return self.loss

    @torch.no_grad()
    def parity_check(self, batch, global_step, rtol=1e-4):
        """Untimed proof for multi-GPU runs (bench.py prints it as `parity_check`): (a) the replicas agree -- float64
        checksums of the sdf grid and both MLPs identical on every rank (bit-for-bit: they see the same reduced gradients),
        k0 to rounding (its row re-scatter is atomic-ordered per rank); (b) the exchanged gradients of one step on this
        rank's `batch` equal the gradients a single GPU computes on the concatenated batch of all ranks.  Leaves gradient
        buffers cleared; call it after the timed region.  -> dict (same on every rank)"""
        import copy
        import torch.distributed as dist
        W, m = self.world, self.m
        self.sync_params()      # (also applies a deferred optimizer phase)
        sums = torch.stack([m.sdf.grid.double().sum(), m.sdf.grid.double().abs().sum(), self.mlp1.flat.double().sum(),
                            self.mlp2.flat.double().sum(), _storage(m.k0.grid).double().sum()])
        allsums = [torch.zeros_like(sums) for _ in range(W)]
        dist.all_gather(allsums, sums)
        identical = all(torch.equal(allsums[0][:4], a[:4]) for a in allsums)
        k0_spread = max(float((a[4] - allsums[0][4]).abs() / allsums[0][4].abs().clamp_min(1e-300)) for a in allsums)
        # (b) one step's gradients through the exchange
        was_eager = self.force_eager
        self.force_eager = True
        self.forward_backward(*batch, global_step)
        self._sync_begin()
        self._sync_k0()
        self._sync_end()
        if self.sharded:      # put the reduced slabs of all ranks together (untimed)
            flat = self.sdf_grad.view(-1)
            dist.all_gather_into_tensor(flat, flat[self.slab[0]:self.slab[1]].clone())
        self.gather_k0_grad()
        self.force_eager = was_eager
        cat = []
        for t in batch:
            parts = [torch.empty_like(t) for _ in range(W)]
            dist.all_gather(parts, t.contiguous())
            cat.append(torch.cat(parts))
        res = torch.zeros(5, dtype=torch.float64, device=self.dev)
        if self.rank == 0:
            m2 = copy.deepcopy(m)
            ref = FusedFineStep(m2, self.N * W, self.cfg, self.rk, row_capacity=self.cap4 * W, world=1, rank=0)
            ref.forward_backward(*cat, global_step)
            pairs = [(self.sdf_grad, ref.sdf_grad), (_storage(self.k0_grad), _storage(ref.k0_grad)),
                     (self.mlp1.flat.grad, ref.mlp1.flat.grad), (self.mlp2.flat.grad, ref.mlp2.flat.grad)]
            ok = 1.0
            for i, (a, b) in enumerate(pairs):
                scale = float(b.abs().max().clamp_min(1e-30))
                d = (a - b).abs()
                bad = d > (rtol * b.abs() + rtol * scale)
                res[i] = float(d.max()) / scale
                # (a handful of ReLU gates of pre-activations within 1e-6 of zero may flip between the two runs: DESIGN.md 6)
                if int(bad.sum()) > max(2, 1e-5 * bad.numel()) or res[i] > 5e-2:
                    ok = 0.0
            res[4] = ok
            del ref, m2
        dist.broadcast(res, 0)
        # leave clean gradient buffers behind
        self.sdf_grad.zero_(); _storage(self.k0_grad).zero_(); self.mlp1.flat.grad.zero_(); self.mlp2.flat.grad.zero_()
        if self.k0_touched is not None:
            self.k0_touched.zero_()
        r = [float(v) for v in res.cpu()]
        return {'ok': bool(identical and r[4] == 1.0 and k0_spread < 1e-6), 'replicas_identical_sdf_mlps': bool(identical),
                'k0_checksum_rel_spread': k0_spread, 'exchange': ('peer-memory pull-reduce + slab Adam with peer stores (sdf, k0), NCCL all-reduce (MLPs)' if self.sdf_peer else
                             'slab reduce-scatter + sharded Adam + param all-gather') if self.sharded else 'dense all-reduce',
                'grad_vs_single_gpu_max_err_over_max': {'sdf': r[0], 'k0': r[1], 'rgbnet': r[2], 'k_rgbnet': r[3]},
                'tolerance': f'|d| <= {rtol} |g| + {rtol} max|g| (all but <= 1e-5 of the elements: ReLU-gate flips)', 'world': W}

    def release_graphs(self):
        """Drop the captured CUDA graphs (they hold NCCL work at world > 1: release them before the process group)."""
        self._graphs.clear()
        self._graph_launches.clear()
        self._render_graph = None
        torch.cuda.synchronize()

    def sync_s_val(self):
        """Write the host-side NeuS s_val into the model's s_val parameter (graph replays keep it on the host only)."""
        if hasattr(self.m, '_s_val_host'):
            self.m.s_val.data.fill_(self.m._s_val_host)

    def _overflow_error(self, n):
        return RuntimeError(f'FusedFineStep: {n} MLP rows exceeded row_capacity={self.cap4} in an earlier step (rows were '
                            'dropped: loss and gradients of that step are wrong); call calibrate() with a representative '
                            'batch at the current global_step, or raise row_capacity')

    def poll_overflow(self, force=False):
        """Raise if an earlier step dropped MLP rows (M4 > row_capacity: its loss and gradients were wrong).  Without
        `force` this never waits for the GPU: see overflow_check_every.  force=True syncs (end of training / of a view)."""
        if force:
            self._ovf_pending = False
            n = int(self.overflow.item())
            if n > 0:
                raise self._overflow_error(n)
            return
        if self._ovf_pending and self._ovf_event.query():
            self._ovf_pending = False
            if int(self._ovf_host[0]) > 0:
                raise self._overflow_error(int(self._ovf_host[0]))
        self._ovf_calls += 1
        if not self._ovf_pending and self._ovf_calls % self.overflow_check_every == 0:
            self._ovf_host.copy_(self.overflow, non_blocking=True)
            self._ovf_event.record()
            self._ovf_pending = True

    def step(self, rays_o, rays_d, viewdirs, target, global_step, grad_sync=None):
        """One training iteration (run.py:600-659).  grad_sync: optional callable run between backward and TV/Adam.
        calibrate() sizes the MLP row buffers; an overflow of that capacity is detected asynchronously (poll_overflow)."""
        loss = self._step(rays_o, rays_d, viewdirs, target, global_step, grad_sync)
        self.poll_overflow()
        return loss

    def _step(self, rays_o, rays_d, viewdirs, target, global_step, grad_sync=None):
        if self.use_graph and grad_sync is None and self.timings is None and self.bitmap_probe is None and not self.force_eager:
            return self._step_graph(rays_o, rays_d, viewdirs, target, global_step)
        if grad_sync is None:
            return self._step_body(rays_o, rays_d, viewdirs, target, global_step, self.tv_flags(global_step))
        loss = self.forward_backward(rays_o, rays_d, viewdirs, target, global_step)
        if grad_sync is not None:
            grad_sync()
        self.regularise(global_step)
        self.optimizer_step()
        return loss

    def counts(self):
        """(M0, M2, M4) of the last step -- this syncs; for logging and capacity calibration only."""
        v = torch.stack([self.offsets[self.N], self.keep_off[self.N].long(), self.off4[self.N].long(), self.overflow[0].long()]).cpu()
        if int(v[3]) > 0:
            raise RuntimeError(f'FusedFineStep: {int(v[3])} MLP rows exceed row_capacity={self.cap4}; call calibrate() or raise it')
        return int(v[0]), int(v[1]), int(v[2])

    def calibrate(self, rays_o, rays_d, viewdirs, global_step=None, headroom=1.3, multiple=4096):
        """Size the row buffers from one forward of a representative batch (one-off sync).  Pass the global_step the
        run is at: the NeuS sharpness s_val (and with it the number of MLP rows) depends on it."""
        self.overflow.zero_()
        self.flush()
        with torch.no_grad():
            self._forward(rays_o, rays_d, viewdirs, global_step, train=False)
        v = torch.stack([self.off4[self.N], self.overflow[0]]).cpu()
        m4 = max(int(v[0]), int(v[1]))
        cap = max(multiple, int(math.ceil(m4 * headroom / multiple)) * multiple)
        self.overflow.zero_()
        if cap != self.cap4:
            self._alloc_rows(cap)
        return cap
