// Per-voxel Adam.
//
// Two semantics are provided because the reference has two:
//  (1) vx_adam_upd -- the reference's CUDA extension (lib/cuda/adam_upd_kernel.cu:9-132, host wrappers
//      :60-132): step_size = lr*sqrt(1-b2^t)/(1-b1^t) folded on the host in float, eps added to the RAW
//      sqrt(v); variants dense / skip-zero-grad / per-voxel lr.
//  (2) vx_adam_step -- the optimizer the reference trainer actually runs, lib/utils.py:154-199: dense,
//      p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).
// Both are one streaming pass: read p,g,m,v once, write p,m,v once (28 B/element; 32 B/element with
// the fused gradient zero-fill that replaces the reference's separate zero-allocation of .grad).
// 128-bit accesses, grid sized as a multiple of the SM count, grid-stride loop.
#include <cstring>
#include "common.cuh"

struct AdamCoef {
  float beta1, beta2, omb1, omb2, eps;
  float step_size;    // (1): lr*sqrt(bc2)/bc1   (2): lr/bc1
  float sqrt_bc2;     // (2) only
};

// Replicas of the parameter array on other GPUs (NVLink peer memory): the owner of a voxel computes its update and stores
// the new parameters into every replica as well -- the "all-gather" of the updated voxels is fused into the optimizer
// pass, moves exactly the bytes that changed, and needs no staging buffer, count or capacity.
#define VX_MAX_PEERS 15
struct VxPeers {
  float* p[VX_MAX_PEERS];
  int n;
};

// mode: 0 dense, 1 skip where grad == 0, 2 per-voxel lr.  kRef = reference-CUDA semantics (1) else (2).
template <bool kRef>
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float perlr, const AdamCoef& c) {
  if (kRef) {
    m = c.beta1 * m + c.omb1 * g;
    v = c.beta2 * v + c.omb2 * g * g;
    p -= c.step_size * perlr * m / (sqrt(v) + c.eps);
  } else {
    // exp_avg.mul_(b1).add_(g, alpha=1-b1); exp_avg_sq.mul_(b2).addcmul_(g, g, value=1-b2)
    m = __fmaf_rn(c.omb1, g, __fmul_rn(m, c.beta1));
    v = __fmaf_rn(__fmul_rn(c.omb2, g), g, __fmul_rn(v, c.beta2));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), c.sqrt_bc2), c.eps);
    p = __fmaf_rn(-c.step_size, __fdiv_rn(__fmul_rn(m, perlr), denom), p);
  }
}

// Sparse-aware dense Adam.  Two optional bitmaps, one bit per group of `group` consecutive elements (a voxel's
// channel vector in the channels-last feature grid):
//   touched : set by the scatter kernels for every voxel that received a gradient THIS step.  Where it is clear the
//             gradient is exactly zero, so it is neither read nor re-zeroed.
//   live    : set for every voxel that has EVER received a gradient (the caller ORs `touched` into it after each
//             step, vx_bitmap_merge).  Where both are clear, exp_avg = exp_avg_sq = 0 and grad = 0, and the dense
//             update of lib/utils.py:154-199 is the identity (m' = 0, v' = 0, p' = p - step * 0 / eps = p): the voxel
//             is skipped without touching HBM.
// The result is bit-identical to the dense pass; only the traffic changes (32 B/element for touched voxels, 24 for
// live-but-untouched, 0 for the rest -- in the fine stage >90 % of a 256^3 grid lies outside the surface shell).
template <bool kRef, int kMode, bool kZeroGrad>
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ param, float* __restrict__ grad,
                                              float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                              const float* __restrict__ perlr, int64_t N, AdamCoef c,
                                              const uint32_t* __restrict__ touched, const uint32_t* __restrict__ live,
                                              uint32_t group, const float* __restrict__ step_dev) {
  if (step_dev) { c.step_size = __ldg(step_dev); c.sqrt_bc2 = __ldg(step_dev + 1); }   // CUDA-graph replays
  const int64_t n4 = N >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float4* p4 = reinterpret_cast<float4*>(param);
  float4* g4 = reinterpret_cast<float4*>(grad);
  float4* m4 = reinterpret_cast<float4*>(exp_avg);
  float4* v4 = reinterpret_cast<float4*>(exp_avg_sq);
  const float4* l4 = reinterpret_cast<const float4*>(perlr);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    bool has_grad = true;
    if (touched) {   // numel < 2^32 checked by the launcher
      const uint32_t e = (uint32_t)i << 2, v0 = e / group, v1 = (e + 3u) / group;
      uint32_t t = (__ldg(touched + (v0 >> 5)) >> (v0 & 31)) & 1u;
      uint32_t l = live ? (__ldg(live + (v0 >> 5)) >> (v0 & 31)) & 1u : 1u;
      if (v1 != v0) {
        t |= (__ldg(touched + (v1 >> 5)) >> (v1 & 31)) & 1u;
        if (live) l |= (__ldg(live + (v1 >> 5)) >> (v1 & 31)) & 1u;
      }
      if (!(t | l)) continue;
      has_grad = t;
    }
    const float4 g = has_grad ? g4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (kMode == 1 && g.x == 0 && g.y == 0 && g.z == 0 && g.w == 0) continue;
    float4 p = p4[i], m = m4[i], v = v4[i];
    float4 l = make_float4(1.f, 1.f, 1.f, 1.f);
    if (kMode == 2) l = l4[i];
    if (kMode != 1 || g.x != 0) adam_one<kRef>(p.x, g.x, m.x, v.x, l.x, c);
    if (kMode != 1 || g.y != 0) adam_one<kRef>(p.y, g.y, m.y, v.y, l.y, c);
    if (kMode != 1 || g.z != 0) adam_one<kRef>(p.z, g.z, m.z, v.z, l.z, c);
    if (kMode != 1 || g.w != 0) adam_one<kRef>(p.w, g.w, m.w, v.w, l.w, c);
    p4[i] = p; m4[i] = m; v4[i] = v;
    if (kZeroGrad && has_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // tail (N % 4)
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const float g = grad[i];
    if (kMode == 1 && g == 0) continue;
    float p = param[i], m = exp_avg[i], v = exp_avg_sq[i];
    adam_one<kRef>(p, g, m, v, kMode == 2 ? perlr[i] : 1.f, c);
    param[i] = p; exp_avg[i] = m; exp_avg_sq[i] = v;
    if (kZeroGrad) grad[i] = 0.f;
  }
}

// The sparse-aware pass proper: one warp per bitmap word (32 voxels = 32 * group consecutive floats, a multiple of 4).
// Words with no live and no touched voxel -- the bulk of the grid -- cost one 8-byte broadcast read; the others are
// walked float4 by float4 by the warp's lanes (coalesced), each lane testing the bits of the voxel(s) its float4 covers.
template <bool kZeroGrad>
__global__ void __launch_bounds__(256) k_adam_sparse(float* __restrict__ param, float* __restrict__ grad,
                                                     float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                     int64_t n_words, int64_t n_vox, AdamCoef c,
                                                     const uint32_t* __restrict__ touched,
                                                     const uint32_t* __restrict__ live, uint32_t group,
                                                     const float* __restrict__ step_dev) {
  if (step_dev) { c.step_size = __ldg(step_dev); c.sqrt_bc2 = __ldg(step_dev + 1); }
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4* p4 = reinterpret_cast<float4*>(param);
  float4* g4 = reinterpret_cast<float4*>(grad);
  float4* m4 = reinterpret_cast<float4*>(exp_avg);
  float4* v4 = reinterpret_cast<float4*>(exp_avg_sq);
  const uint32_t per_word4 = 8u * group;       // float4s per bitmap word
  // lanes prefetch 32 words at a time, then the warp walks the non-empty ones
  for (int64_t wbase = warp0 * 32; wbase < n_words; wbase += n_warps * 32) {
    const int64_t wi = wbase + lane;
    const uint32_t tw_l = wi < n_words ? __ldg(touched + wi) : 0u;
    const uint32_t lw_l = (wi < n_words && live) ? __ldg(live + wi) : (wi < n_words && !live ? 0xffffffffu : 0u);
    uint32_t busy = __ballot_sync(0xffffffffu, (tw_l | lw_l) != 0u);
    while (busy) {
      const int src = __ffs(busy) - 1;
      busy &= busy - 1;
      const uint32_t tw = __shfl_sync(0xffffffffu, tw_l, src), lw = __shfl_sync(0xffffffffu, lw_l, src);
      const int64_t w = wbase + src;
      const int64_t f4_base = w * per_word4;
      const int64_t f4_end = min(f4_base + per_word4, (n_vox * group) >> 2);
      for (int64_t i = f4_base + lane; i < f4_end; i += 32) {
        const uint32_t e = (uint32_t)(i - f4_base) << 2, b0 = e / group, b1 = min((e + 3u) / group, 31u);
        const uint32_t t = ((tw >> b0) | (tw >> b1)) & 1u, l = ((lw >> b0) | (lw >> b1)) & 1u;
        if (!(t | l)) continue;
        const float4 g = t ? g4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 p = p4[i], m = m4[i], v = v4[i];
        adam_one<false>(p.x, g.x, m.x, v.x, 1.f, c);
        adam_one<false>(p.y, g.y, m.y, v.y, 1.f, c);
        adam_one<false>(p.z, g.z, m.z, v.z, 1.f, c);
        adam_one<false>(p.w, g.w, m.w, v.w, 1.f, c);
        p4[i] = p; m4[i] = m; v4[i] = v;
        if (kZeroGrad && t) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

template <bool kRef>
static int launch_adam(float* param, float* grad, float* m, float* v, const float* perlr, int64_t N, const AdamCoef& c,
                       int mode, int zero_grad, cudaStream_t st, const uint32_t* touched = nullptr,
                       const uint32_t* live = nullptr, int group = 1, const float* step_dev = nullptr) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                         reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) |
                         reinterpret_cast<uintptr_t>(perlr)) & 15) == 0;
  if (!aligned) { vx_set_error("vx_adam", "tensors must be 16-byte aligned"); return -1; }
  const int64_t want = (N / 4 + 255) / 256 + 1;
  const int blocks = (int)min(want, (int64_t)vx_num_sms() * 8);
#define VX_ADAM(MODE, ZG) k_adam<kRef, MODE, ZG><<<blocks, 256, 0, st>>>(param, grad, m, v, perlr, N, c, touched, live, (uint32_t)group, step_dev)
  if (mode == 0) { if (zero_grad) VX_ADAM(0, true); else VX_ADAM(0, false); }
  else if (mode == 1) { if (zero_grad) VX_ADAM(1, true); else VX_ADAM(1, false); }
  else { if (zero_grad) VX_ADAM(2, true); else VX_ADAM(2, false); }
#undef VX_ADAM
  return vx_check_launch("vx_adam");
}

VX_API int vx_adam_upd(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const float* perlr, int64_t N,
                       int step, float beta1, float beta2, float lr, float eps, int mode, cudaStream_t st) {
  if (N <= 0) return 0;
  VX_REQUIRE(mode >= 0 && mode <= 2, "vx_adam_upd", "mode must be 0, 1 or 2");
  VX_REQUIRE(mode != 2 || perlr, "vx_adam_upd", "mode 2 needs perlr");
  AdamCoef c;
  c.beta1 = beta1; c.beta2 = beta2; c.omb1 = 1 - beta1; c.omb2 = 1 - beta2; c.eps = eps; c.sqrt_bc2 = 1.f;
  c.step_size = lr * sqrtf(1 - powf(beta2, (float)step)) / (1 - powf(beta1, (float)step));  // adam_upd_kernel.cu:72
  return launch_adam<true>(param, const_cast<float*>(grad), exp_avg, exp_avg_sq, perlr, N, c, mode, 0, st);
}

// Block-sparse dense Adam for a single-channel grid (the sdf grid): one warp per block of 128 consecutive elements.  The
// warp reads the block's gradient (the one compulsory pass over the grid: 4 B / element); if the whole block has a zero
// gradient and has never had a non-zero one (`live_blocks[b] == 0`: its moments are still exactly zero), the dense update
// of lib/utils.py:154-199 is the identity there (m' = 0, v' = 0, p' = p - step * 0 / eps = p) and the block is skipped:
// no parameter / moment traffic.  Otherwise the block is updated exactly like k_adam does (and flagged live for good).
// Bit-identical to the dense pass.  In the fine stage the ray gradients and the TV gradients both live in the shell around
// the surface / inside the non-empty mask: ~80 % of a 256^3 grid is skipped (537 MB -> ~170 MB per step).
template <bool kZeroGrad>
__global__ void __launch_bounds__(256) k_adam_blocklive(float* __restrict__ param, float* __restrict__ grad,
                                                        float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                        int64_t n_blocks, AdamCoef c, uint8_t* __restrict__ live_blocks,
                                                        const float* __restrict__ step_dev, const VxPeers peers) {
  if (step_dev) { c.step_size = __ldg(step_dev); c.sqrt_bc2 = __ldg(step_dev + 1); }
  const int lane = threadIdx.x & 31;
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float4* p4 = reinterpret_cast<float4*>(param);
  float4* g4 = reinterpret_cast<float4*>(grad);
  float4* m4 = reinterpret_cast<float4*>(exp_avg);
  float4* v4 = reinterpret_cast<float4*>(exp_avg_sq);
  for (int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < n_blocks; b += n_warps) {
    const int64_t i = b * 32 + lane;
    const float4 g = g4[i];
    const bool nz = (g.x != 0.f) | (g.y != 0.f) | (g.z != 0.f) | (g.w != 0.f);
    const bool any = __any_sync(0xffffffffu, nz);
    const bool lv = live_blocks[b] != 0;
    if (!any && !lv) continue;
    float4 p = p4[i], m = m4[i], v = v4[i];
    adam_one<false>(p.x, g.x, m.x, v.x, 1.f, c);
    adam_one<false>(p.y, g.y, m.y, v.y, 1.f, c);
    adam_one<false>(p.z, g.z, m.z, v.z, 1.f, c);
    adam_one<false>(p.w, g.w, m.w, v.w, 1.f, c);
    p4[i] = p; m4[i] = m; v4[i] = v;
    for (int r = 0; r < peers.n; ++r) reinterpret_cast<float4*>(peers.p[r])[i] = p;     // the replicas on the other GPUs
    if (kZeroGrad && nz) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!lv && lane == 0) live_blocks[b] = 1;
  }
}

// trainer semantics (lib/utils.py:154-199), numel % 128 == 0, 16-byte aligned tensors; live_blocks: numel / 128 bytes,
// zero-initialised by the caller once (all ones after loading moments from elsewhere)
static int adam_step_blocklive(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t N, float beta1,
                               float beta2, float one_minus_beta1, float one_minus_beta2, float step_size,
                               float sqrt_bias_correction2, float eps, int zero_grad, uint8_t* live_blocks,
                               const float* step_dev, const VxPeers& peers, cudaStream_t st) {
  if (N <= 0) return 0;
  const bool aligned = ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                         reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0;
  VX_REQUIRE(aligned && N % 128 == 0 && live_blocks, "vx_adam_step_blocklive", "16-byte aligned tensors, numel % 128 == 0, live_blocks");
  AdamCoef c;
  c.beta1 = beta1; c.beta2 = beta2; c.omb1 = one_minus_beta1; c.omb2 = one_minus_beta2; c.eps = eps;
  c.step_size = step_size; c.sqrt_bc2 = sqrt_bias_correction2;
  const int64_t n_blocks = N / 128;
  const int blocks = (int)min((n_blocks + 7) / 8, (int64_t)vx_num_sms() * 8);
  if (zero_grad) k_adam_blocklive<true><<<blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n_blocks, c, live_blocks, step_dev, peers);
  else k_adam_blocklive<false><<<blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n_blocks, c, live_blocks, step_dev, peers);
  return vx_check_launch("vx_adam_step_blocklive");
}

static int make_peers(const uint64_t* peer_ptrs_host, int n_peers, VxPeers& peers, const char* where) {
  VX_REQUIRE(n_peers >= 0 && n_peers <= VX_MAX_PEERS && (n_peers == 0 || peer_ptrs_host), where, "0 <= n_peers <= 15");
  peers.n = n_peers;
  for (int r = 0; r < n_peers; ++r) {
    VX_REQUIRE((peer_ptrs_host[r] & 15) == 0, where, "peer arrays must be 16-byte aligned");
    peers.p[r] = reinterpret_cast<float*>(peer_ptrs_host[r]);
  }
  return 0;
}

VX_API int vx_adam_step_blocklive(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t N, float beta1,
                                  float beta2, float one_minus_beta1, float one_minus_beta2, float step_size,
                                  float sqrt_bias_correction2, float eps, int zero_grad, uint8_t* live_blocks,
                                  const float* step_dev, cudaStream_t st) {
  VxPeers peers;
  peers.n = 0;
  return adam_step_blocklive(param, grad, exp_avg, exp_avg_sq, N, beta1, beta2, one_minus_beta1, one_minus_beta2, step_size,
                             sqrt_bias_correction2, eps, zero_grad, live_blocks, step_dev, peers, st);
}

// vx_adam_step_blocklive on the slab of a replicated single-channel grid this rank owns, storing the parameters of every
// updated block into the same slab of the replicas on n_peers other GPUs (see vx_adam_step_worklist_peers)
VX_API int vx_adam_step_blocklive_peers(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t N, float beta1,
                                        float beta2, float one_minus_beta1, float one_minus_beta2, float step_size,
                                        float sqrt_bias_correction2, float eps, int zero_grad, uint8_t* live_blocks,
                                        const float* step_dev, const uint64_t* peer_params_host, int n_peers,
                                        cudaStream_t st) {
  VxPeers peers;
  if (int rc = make_peers(peer_params_host, n_peers, peers, "vx_adam_step_blocklive_peers")) return rc;
  return adam_step_blocklive(param, grad, exp_avg, exp_avg_sq, N, beta1, beta2, one_minus_beta1, one_minus_beta2, step_size,
                             sqrt_bias_correction2, eps, zero_grad, live_blocks, step_dev, peers, st);
}

// ---------------------------------------------------------------------------------------------
// Sparse reduce-scatter of a single-channel gradient grid over peer memory.  The ray gradients of the sdf grid live in
// the shell around the surface (~10 % of the 128-voxel blocks): instead of an NCCL reduce-scatter of the dense 64 MB grid,
// every rank flags its non-zero blocks (vx_block_nonzero), and after a cross-rank barrier the owner of a slab PULLS, block
// by block, only the flagged blocks of every rank's gradient (its own included, in rank order: a fixed summation order)
// over NVLink and leaves scale * sum in its own slab.  The other ranks' copies of the slab are stale afterwards and are
// cleared by their holders behind a second barrier.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_block_nonzero(const float* __restrict__ g, int64_t n_blocks, uint8_t* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < n_blocks; b += n_warps) {
    const float4 v = g4[b * 32 + lane];
    const bool nz = (v.x != 0.f) | (v.y != 0.f) | (v.z != 0.f) | (v.w != 0.f);
    const bool any = __any_sync(0xffffffffu, nz);
    if (lane == 0) mask[b] = any ? 1 : 0;
  }
}

VX_API int vx_block_nonzero(const float* g, int64_t N, uint8_t* mask, cudaStream_t st) {
  if (N <= 0) return 0;
  VX_REQUIRE(N % 128 == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0, "vx_block_nonzero", "numel % 128 == 0, 16-byte aligned");
  const int64_t n_blocks = N / 128;
  k_block_nonzero<<<(int)min((n_blocks + 7) / 8, (int64_t)vx_num_sms() * 8), 256, 0, st>>>(g, n_blocks, mask);
  return vx_check_launch("vx_block_nonzero");
}

struct VxRanks {
  const float* g[VX_MAX_PEERS + 1];
  const uint8_t* m[VX_MAX_PEERS + 1];
  int n;
};

__global__ void __launch_bounds__(256) k_pull_reduce(float* out, const VxRanks ranks, int64_t n_blocks, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < n_blocks; b += n_warps) {
    uint32_t has = 0u;
    if (lane < ranks.n) has = ranks.m[lane][b];          // one flag per rank, fetched in parallel
    const uint32_t bits = __ballot_sync(0xffffffffu, has != 0u);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v[VX_MAX_PEERS + 1];
#pragma unroll
    for (int q = 0; q <= VX_MAX_PEERS; ++q)             // issue every load before the first use
      if (q < ranks.n && ((bits >> q) & 1u)) v[q] = reinterpret_cast<const float4*>(ranks.g[q])[b * 32 + lane];
#pragma unroll
    for (int q = 0; q <= VX_MAX_PEERS; ++q)
      if (q < ranks.n && ((bits >> q) & 1u)) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
    if (bits) reinterpret_cast<float4*>(out)[b * 32 + lane] = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
  }
}

// out (this rank's own slab, N elements; may alias grads_host[own rank]) = scale * sum over the n_ranks arrays (device addresses
// of the same slab on every rank, own included, in rank order) of the blocks their masks flag; blocks no rank flags keep
// their (zero) content.
VX_API int vx_pull_reduce(float* out, int64_t N, const uint64_t* grads_host, const uint64_t* masks_host, int n_ranks,
                          float scale, cudaStream_t st) {
  if (N <= 0) return 0;
  VX_REQUIRE(N % 128 == 0 && n_ranks >= 1 && n_ranks <= VX_MAX_PEERS + 1 && grads_host && masks_host, "vx_pull_reduce",
             "numel % 128 == 0, 1 <= n_ranks <= 16");
  VxRanks r;
  r.n = n_ranks;
  for (int q = 0; q < n_ranks; ++q) {
    VX_REQUIRE((grads_host[q] & 15) == 0, "vx_pull_reduce", "16-byte aligned arrays");
    r.g[q] = reinterpret_cast<const float*>(grads_host[q]);
    r.m[q] = reinterpret_cast<const uint8_t*>(masks_host[q]);
  }
  const int64_t n_blocks = N / 128;
  k_pull_reduce<<<(int)min((n_blocks + 7) / 8, (int64_t)vx_num_sms() * 8), 256, 0, st>>>(out, r, n_blocks, scale);
  return vx_check_launch("vx_pull_reduce");
}

// bias corrections are computed by the caller in Python doubles exactly like lib/utils.py:176-177,192
VX_API int vx_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, const float* perlr, int64_t N,
                        float beta1, float beta2, float one_minus_beta1, float one_minus_beta2, float step_size,
                        float sqrt_bias_correction2, float eps, int skip_zero_grad, int zero_grad,
                        const uint32_t* touched, const uint32_t* live, int group, const float* step_dev, cudaStream_t st) {
  if (N <= 0) return 0;
  // a float4 of 4 consecutive elements spans at most two voxels only when a voxel has >= 3 elements: the bitmap tests
  // look at the voxels of its first and last element
  VX_REQUIRE(!touched || (group >= 3 && N % 4 == 0 && N < ((int64_t)1 << 32)), "vx_adam_step",
             "touched bitmap needs group >= 3, numel % 4 == 0 and numel < 2^32");
  VX_REQUIRE(!live || touched, "vx_adam_step", "a live bitmap needs a touched bitmap");
  AdamCoef c;
  c.beta1 = beta1; c.beta2 = beta2; c.omb1 = one_minus_beta1; c.omb2 = one_minus_beta2; c.eps = eps;
  c.step_size = step_size; c.sqrt_bc2 = sqrt_bias_correction2;
  const int mode = perlr ? 2 : (skip_zero_grad ? 1 : 0);
  if (touched && mode == 0) {
    const bool aligned = ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                           reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0;
    VX_REQUIRE(aligned && N % group == 0, "vx_adam_step", "bitmap pass: 16-byte aligned tensors, numel % group == 0");
    const int64_t n_vox = N / group, n_words = (n_vox + 31) / 32;
    const int blocks = (int)min((n_words + 255) / 256, (int64_t)vx_num_sms() * 8);
    if (zero_grad) k_adam_sparse<true><<<blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n_words, n_vox, c, touched, live, (uint32_t)group, step_dev);
    else k_adam_sparse<false><<<blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n_words, n_vox, c, touched, live, (uint32_t)group, step_dev);
    return vx_check_launch("vx_adam_step");
  }
  return launch_adam<false>(param, grad, exp_avg, exp_avg_sq, perlr, N, c, mode, zero_grad, st, touched, live, group, step_dev);
}

// ---------------------------------------------------------------------------------------------
// Work-list form of the sparse-aware pass.  k_adam_sparse deals the bitmap words to the warps statically; the live voxels
// sit in the shell around the surface, so a few warps own most of them and walk them one dependent round trip after the
// other, most lanes idle (a word holds 1-3 live voxels: 112 us for 158 MB).  Here a first kernel expands the bitmaps into a
// list of live voxels (warp-scanned counts, one atomic per warp; the order is irrelevant, every voxel is independent; the
// touched flag rides in bit 31) and performs vx_bitmap_merge (live |= touched, touched = 0) on the way; a second kernel
// deals the LIST to the threads, one 16-byte (8- / 4-byte for channel counts that are not multiples of 4 / 2) piece of a
// voxel per thread: every lane busy, every load independent.  Same arithmetic per element: bit-identical results.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bitmap_voxel_list(uint32_t* __restrict__ touched, uint32_t* __restrict__ live,
                                                           int64_t n_words, int64_t n_vox, int merge,
                                                           uint32_t* __restrict__ list, uint32_t* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_round = (n_words + 31) & ~(int64_t)31;      // whole warps stay converged for the scan
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_round; w += stride) {
    uint32_t tw = 0u, lw = 0u;
    if (w < n_words) {
      tw = touched[w];
      lw = live ? live[w] : 0xffffffffu;
      if (w == n_words - 1 && (n_vox & 31)) {          // bits beyond the last voxel do not exist
        const uint32_t valid = (1u << (n_vox & 31)) - 1u;
        tw &= valid; lw &= valid;
      }
    }
    const uint32_t busy = tw | lw;
    if (!__ballot_sync(0xffffffffu, busy != 0u)) continue;
    const int n = __popc(busy);
    int incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += y;
    }
    uint32_t base = 0;
    if (lane == 31) base = atomicAdd(count, (uint32_t)incl);
    base = __shfl_sync(0xffffffffu, base, 31) + (uint32_t)(incl - n);
    uint32_t b = busy;
    while (b) {
      const int bit = __ffs(b) - 1;
      b &= b - 1;
      list[base++] = ((uint32_t)w * 32u + (uint32_t)bit) | (((tw >> bit) & 1u) << 31);
    }
    if (merge && tw) { if (live) live[w] = lw | tw; touched[w] = 0u; }
  }
}

template <int kW> struct VxVec;
template <> struct VxVec<4> { typedef float4 T; };
template <> struct VxVec<2> { typedef float2 T; };
template <> struct VxVec<1> { typedef float T; };

template <int kW, bool kZeroGrad>
__global__ void __launch_bounds__(256) k_adam_voxel_list(float* __restrict__ param, float* __restrict__ grad,
                                                         float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                         AdamCoef c, uint32_t parts, const uint32_t* __restrict__ list,
                                                         const uint32_t* __restrict__ count,
                                                         const float* __restrict__ step_dev, const VxPeers peers) {
  typedef typename VxVec<kW>::T V;
  if (step_dev) { c.step_size = __ldg(step_dev); c.sqrt_bc2 = __ldg(step_dev + 1); }
  const int64_t n = (int64_t)(*count) * parts;
  V* p4 = reinterpret_cast<V*>(param);
  V* g4 = reinterpret_cast<V*>(grad);
  V* m4 = reinterpret_cast<V*>(exp_avg);
  V* v4 = reinterpret_cast<V*>(exp_avg_sq);
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t e = __ldg(list + k / parts);
    const bool t = (e >> 31) != 0u;
    const int64_t i = (int64_t)(e & 0x7fffffffu) * parts + (k % parts);
    V g;
    float* gf = reinterpret_cast<float*>(&g);
    if (t) g = g4[i];
    else {
#pragma unroll
      for (int j = 0; j < kW; ++j) gf[j] = 0.f;
    }
    V p = p4[i], m = m4[i], v = v4[i];
    float* pf = reinterpret_cast<float*>(&p);
    float* mf = reinterpret_cast<float*>(&m);
    float* vf = reinterpret_cast<float*>(&v);
#pragma unroll
    for (int j = 0; j < kW; ++j) adam_one<false>(pf[j], gf[j], mf[j], vf[j], 1.f, c);
    p4[i] = p; m4[i] = m; v4[i] = v;
    for (int r = 0; r < peers.n; ++r) reinterpret_cast<V*>(peers.p[r])[i] = p;
    if (kZeroGrad && t) {
#pragma unroll
      for (int j = 0; j < kW; ++j) gf[j] = 0.f;
      g4[i] = g;
    }
  }
}

// trainer semantics (lib/utils.py:154-199) over the voxels flagged in touched | live (see k_adam above), work-list form.
// group: elements per voxel; work: numel / group + 1 uint32 of scratch (voxel list + its counter at the end);
// merge != 0: live |= touched, touched = 0 in the same pass (then vx_bitmap_merge is not needed).
static int adam_step_worklist(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t N, float beta1,
                              float beta2, float one_minus_beta1, float one_minus_beta2, float step_size,
                              float sqrt_bias_correction2, float eps, int zero_grad, uint32_t* touched, uint32_t* live,
                              int group, int merge, uint32_t* work, const float* step_dev, const VxPeers& peers,
                              cudaStream_t st) {
  if (N <= 0) return 0;
  const bool aligned = ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                         reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0;
  VX_REQUIRE(touched && work && aligned && group >= 1 && N % group == 0 && N / group < ((int64_t)1 << 31),
             "vx_adam_step_worklist", "needs touched, work, 16-byte aligned tensors, numel % group == 0, numel / group < 2^31");
  AdamCoef c;
  c.beta1 = beta1; c.beta2 = beta2; c.omb1 = one_minus_beta1; c.omb2 = one_minus_beta2; c.eps = eps;
  c.step_size = step_size; c.sqrt_bc2 = sqrt_bias_correction2;
  const int64_t n_vox = N / group, n_words = (n_vox + 31) / 32;
  uint32_t* count = work + n_vox;
  cudaError_t e = cudaMemsetAsync(count, 0, sizeof(uint32_t), st);
  if (e != cudaSuccess) { vx_set_error("vx_adam_step_worklist", cudaGetErrorString(e)); return -1; }
  k_bitmap_voxel_list<<<(int)min((n_words + 255) / 256, (int64_t)vx_num_sms() * 8), 256, 0, st>>>(touched, live, n_words, n_vox, merge, work, count);
  int rc = vx_check_launch("vx_adam_step_worklist(list)");
  if (rc) return rc;
  const int blocks = vx_num_sms() * 8;
  const int w = group % 4 == 0 ? 4 : (group % 2 == 0 ? 2 : 1);
  const uint32_t parts = (uint32_t)(group / w);
#define VX_AVL(W, ZG) k_adam_voxel_list<W, ZG><<<blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, c, parts, work, count, step_dev, peers)
  if (w == 4) { if (zero_grad) VX_AVL(4, true); else VX_AVL(4, false); }
  else if (w == 2) { if (zero_grad) VX_AVL(2, true); else VX_AVL(2, false); }
  else { if (zero_grad) VX_AVL(1, true); else VX_AVL(1, false); }
#undef VX_AVL
  return vx_check_launch("vx_adam_step_worklist");
}

VX_API int vx_adam_step_worklist(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t N, float beta1,
                                 float beta2, float one_minus_beta1, float one_minus_beta2, float step_size,
                                 float sqrt_bias_correction2, float eps, int zero_grad, uint32_t* touched, uint32_t* live,
                                 int group, int merge, uint32_t* work, const float* step_dev, cudaStream_t st) {
  VxPeers peers;
  peers.n = 0;
  return adam_step_worklist(param, grad, exp_avg, exp_avg_sq, N, beta1, beta2, one_minus_beta1, one_minus_beta2, step_size,
                            sqrt_bias_correction2, eps, zero_grad, touched, live, group, merge, work, step_dev, peers, st);
}

// vx_adam_step_worklist on the slice of a replicated parameter array this rank owns; the updated parameters are also stored
// into the same slice of the replicas on n_peers other GPUs (peer_params_host: their device addresses, peer access enabled --
// vx_enable_peer_access).  Peer stores are complete when the kernel is; order them against the readers with a cross-rank
// barrier on the same stream.
VX_API int vx_adam_step_worklist_peers(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t N, float beta1,
                                       float beta2, float one_minus_beta1, float one_minus_beta2, float step_size,
                                       float sqrt_bias_correction2, float eps, int zero_grad, uint32_t* touched,
                                       uint32_t* live, int group, int merge, uint32_t* work, const float* step_dev,
                                       const uint64_t* peer_params_host, int n_peers, cudaStream_t st) {
  VxPeers peers;
  if (int rc = make_peers(peer_params_host, n_peers, peers, "vx_adam_step_worklist_peers")) return rc;
  return adam_step_worklist(param, grad, exp_avg, exp_avg_sq, N, beta1, beta2, one_minus_beta1, one_minus_beta2, step_size,
                            sqrt_bias_correction2, eps, zero_grad, touched, live, group, merge, work, step_dev, peers, st);
}

// CUDA IPC for the replicas of a parameter array held by the other ranks of the node (one process per GPU): the owner exports
// the handle of the cudaMalloc allocation (base pointer), a peer opens it with ITS OWN device current -- the mapping is then
// addressable by that device's kernels over NVLink (cudaIpcMemLazyEnablePeerAccess).
VX_API int vx_ipc_get_handle(const void* base_ptr, uint8_t* handle_host /* 64 bytes */) {
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(base_ptr));
  if (e != cudaSuccess) { vx_set_error("vx_ipc_get_handle", cudaGetErrorString(e)); cudaGetLastError(); return -1; }
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_host, &h, sizeof(h));
  return 0;
}
VX_API int vx_ipc_open_handle(const uint8_t* handle_host, uint64_t* ptr_out_host) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { vx_set_error("vx_ipc_open_handle", cudaGetErrorString(e)); cudaGetLastError(); return -1; }
  *ptr_out_host = reinterpret_cast<uint64_t>(p);
  return 0;
}
VX_API int vx_ipc_close_handle(uint64_t ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(reinterpret_cast<void*>(ptr));
  if (e != cudaSuccess) { vx_set_error("vx_ipc_close_handle", cudaGetErrorString(e)); cudaGetLastError(); return -1; }
  return 0;
}

// cudaDeviceEnablePeerAccess(peer) for the current device (already enabled = success)
VX_API int vx_enable_peer_access(int peer_device) {
  int cur = -1, can = 0;
  cudaGetDevice(&cur);
  if (cur == peer_device) return 0;
  cudaError_t e = cudaDeviceCanAccessPeer(&can, cur, peer_device);
  if (e != cudaSuccess || !can) { cudaGetLastError(); vx_set_error("vx_enable_peer_access", "no peer access between these devices"); return -1; }
  e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
  if (e != cudaSuccess) { vx_set_error("vx_enable_peer_access", cudaGetErrorString(e)); cudaGetLastError(); return -1; }
  return 0;
}

// live |= touched; touched = 0  (after the Adam pass that consumed both)
__global__ void k_bitmap_merge(uint32_t* __restrict__ live, uint32_t* __restrict__ touched, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t t = touched[i];
    if (t) { live[i] |= t; touched[i] = 0u; }
  }
}

VX_API int vx_bitmap_merge(uint32_t* live, uint32_t* touched, int64_t n_words, cudaStream_t st) {
  if (n_words <= 0) return 0;
  k_bitmap_merge<<<(int)min((n_words + 255) / 256, (int64_t)vx_num_sms() * 8), 256, 0, st>>>(live, touched, n_words);
  return vx_check_launch("vx_bitmap_merge");
}
