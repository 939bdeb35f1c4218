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
#include "common.cuh"

struct AdamCoef {
  float beta1, beta2, omb1, omb2, eps;
  float step_size;    // (1): lr*sqrt(bc2)/bc1   (2): lr/bc1
  float sqrt_bc2;     // (2) only
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

template <bool kRef, int kMode, bool kZeroGrad>
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ param, float* __restrict__ grad,
                                              float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                              const float* __restrict__ perlr, int64_t N, AdamCoef c) {
  const int64_t n4 = N >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float4* p4 = reinterpret_cast<float4*>(param);
  float4* g4 = reinterpret_cast<float4*>(grad);
  float4* m4 = reinterpret_cast<float4*>(exp_avg);
  float4* v4 = reinterpret_cast<float4*>(exp_avg_sq);
  const float4* l4 = reinterpret_cast<const float4*>(perlr);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 g = g4[i];
    if (kMode == 1 && g.x == 0 && g.y == 0 && g.z == 0 && g.w == 0) continue;
    float4 p = p4[i], m = m4[i], v = v4[i];
    float4 l = make_float4(1.f, 1.f, 1.f, 1.f);
    if (kMode == 2) l = l4[i];
    if (kMode != 1 || g.x != 0) adam_one<kRef>(p.x, g.x, m.x, v.x, l.x, c);
    if (kMode != 1 || g.y != 0) adam_one<kRef>(p.y, g.y, m.y, v.y, l.y, c);
    if (kMode != 1 || g.z != 0) adam_one<kRef>(p.z, g.z, m.z, v.z, l.z, c);
    if (kMode != 1 || g.w != 0) adam_one<kRef>(p.w, g.w, m.w, v.w, l.w, c);
    p4[i] = p; m4[i] = m; v4[i] = v;
    if (kZeroGrad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
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

template <bool kRef>
static int launch_adam(float* param, float* grad, float* m, float* v, const float* perlr, int64_t N, const AdamCoef& c,
                       int mode, int zero_grad, cudaStream_t st) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                         reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) |
                         reinterpret_cast<uintptr_t>(perlr)) & 15) == 0;
  if (!aligned) { vx_set_error("vx_adam", "tensors must be 16-byte aligned"); return -1; }
  const int64_t want = (N / 4 + 255) / 256 + 1;
  const int blocks = (int)min(want, (int64_t)vx_num_sms() * 8);
#define VX_ADAM(MODE, ZG) k_adam<kRef, MODE, ZG><<<blocks, 256, 0, st>>>(param, grad, m, v, perlr, N, c)
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

// bias corrections are computed by the caller in Python doubles exactly like lib/utils.py:176-177,192
VX_API int vx_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, const float* perlr, int64_t N,
                        float beta1, float beta2, float one_minus_beta1, float one_minus_beta2, float step_size,
                        float sqrt_bias_correction2, float eps, int skip_zero_grad, int zero_grad, cudaStream_t st) {
  if (N <= 0) return 0;
  AdamCoef c;
  c.beta1 = beta1; c.beta2 = beta2; c.omb1 = one_minus_beta1; c.omb2 = one_minus_beta2; c.eps = eps;
  c.step_size = step_size; c.sqrt_bc2 = sqrt_bias_correction2;
  const int mode = perlr ? 2 : (skip_zero_grad ? 1 : 0);
  return launch_adam<false>(param, grad, exp_avg, exp_avg_sq, perlr, N, c, mode, zero_grad, st);
}
