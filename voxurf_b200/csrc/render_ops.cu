// raw2alpha / alpha2weight.
// Reference: /root/reference/lib/cuda/render_utils_kernel.cu:430-707.
//
// alpha2weight is re-designed: the reference runs ONE THREAD per ray (8192 threads, a serial
// ~200-step loop each, uncoalesced).  Here one warp owns a ray, loads 32 alphas coalesced and
// replays the reference's recurrence  T <- float(double(T) * (1. - alpha))  in lock-step across
// the warp (every lane computes the same chain from shuffled alphas, lane j keeps step j).
// The arithmetic and its order are exactly the reference's, so weights, T, alphainv_last and the
// early-exit index i_end are BIT-EXACT -- no re-association, no flip caveat.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// raw2alpha (render_utils_kernel.cu:430-574)
// ---------------------------------------------------------------------------------------------
__global__ void k_raw2alpha(const float* __restrict__ density, float shift, const float* __restrict__ interval_vec,
                            float interval, int64_t n, float* __restrict__ exp_d, float* __restrict__ alpha) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float iv = interval_vec ? interval_vec[i] : interval;
    const float e = exp(density[i] + shift);  // can be inf
    exp_d[i] = e;
    alpha[i] = 1 - pow(1 + e, -iv);
  }
}

__global__ void k_raw2alpha_backward(const float* __restrict__ exp_d, const float* __restrict__ grad_back,
                                     const float* __restrict__ interval_vec, float interval, int64_t n,
                                     float* __restrict__ grad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float iv = interval_vec ? interval_vec[i] : interval;
    grad[i] = min(exp_d[i], 1e10) * pow(1 + exp_d[i], -iv - 1) * iv * grad_back[i];
  }
}

VX_API int vx_raw2alpha(const float* density, float shift, const float* interval_vec, float interval, int64_t n,
                        float* exp_d, float* alpha, cudaStream_t st) {
  if (n <= 0) return 0;
  k_raw2alpha<<<vx_blocks(n, 256), 256, 0, st>>>(density, shift, interval_vec, interval, n, exp_d, alpha);
  return vx_check_launch("vx_raw2alpha");
}

VX_API int vx_raw2alpha_backward(const float* exp_d, const float* grad_back, const float* interval_vec, float interval,
                                 int64_t n, float* grad, cudaStream_t st) {
  if (n <= 0) return 0;
  k_raw2alpha_backward<<<vx_blocks(n, 256), 256, 0, st>>>(exp_d, grad_back, interval_vec, interval, n, grad);
  return vx_check_launch("vx_raw2alpha_backward");
}

// ---------------------------------------------------------------------------------------------
// alpha2weight (render_utils_kernel.cu:576-651)
// ---------------------------------------------------------------------------------------------
__global__ void k_a2w_init(int n_rays, float* __restrict__ alphainv_last, int64_t* __restrict__ i_start,
                           int64_t* __restrict__ i_end) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rays) { alphainv_last[r] = 1.f; i_start[r] = 0; i_end[r] = 0; }
}

// render_utils_kernel.cu:607-617 + the host-side  i_end[ray_id[n_pts-1]] = n_pts  (:635), done on
// the device here (the reference's host-indexed write forces a sync).
__global__ void k_a2w_segments(const int64_t* __restrict__ ray_id, int64_t n_pts, int64_t* __restrict__ i_start,
                               int64_t* __restrict__ i_end) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pts) {
    const int64_t r = ray_id[i];
    if (i > 0) {
      const int64_t rp = ray_id[i - 1];
      if (r != rp) { i_start[r] = i; i_end[rp] = i; }
    }
    if (i == n_pts - 1) i_end[r] = n_pts;
  }
}

// The warp-cooperative recurrence shared by the legacy and the fused entry points.
// Processes samples [i_s, i_e_max) of one ray; returns the stop index and final T.
template <bool kHasKeep>
__device__ __forceinline__ void a2w_ray(const float* __restrict__ alpha, const uint8_t* __restrict__ keep, int64_t i_s,
                                        int64_t i_e_max, float* __restrict__ weight, float* __restrict__ T,
                                        int64_t& i_stop, float& T_last) {
  const int lane = threadIdx.x & 31;
  float T_cum = 1.;
  bool done = false;
  int64_t stop = i_e_max;
  for (int64_t base = i_s; base < i_e_max; base += 32) {
    const int64_t i = base + lane;
    const bool valid = i < i_e_max;
    const float a = valid ? alpha[i] : 0.f;
    const bool k = kHasKeep ? (valid && keep[i]) : valid;
    const uint32_t kmask = __ballot_sync(0xffffffffu, k);
    float myT = 1.f, myW = 0.f;
    if (!done) {
      const int cnt = (int)min((int64_t)32, i_e_max - base);
      for (int j = 0; j < cnt; ++j) {
        if (kHasKeep && !((kmask >> j) & 1u)) continue;  // sample dropped by a threshold: not on the ray
        const float aj = __shfl_sync(0xffffffffu, a, j);
        if (lane == j) { myT = T_cum; myW = T_cum * aj; }
        T_cum *= (1. - aj);
        if (T_cum < 1e-3) { done = true; stop = base + j + 1; break; }
      }
    }
    if (valid) { T[i] = myT; weight[i] = myW; }
    // after the exit the tail keeps weight = 0, T = 1 (reference pre-fill, :624-625)
  }
  i_stop = stop;
  T_last = T_cum;
}

__global__ void k_alpha2weight(const float* __restrict__ alpha, int n_rays, float* __restrict__ weight,
                               float* __restrict__ T, float* __restrict__ alphainv_last,
                               const int64_t* __restrict__ i_start, int64_t* __restrict__ i_end) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * warps_per_block) {
    int64_t stop;
    float T_last;
    a2w_ray<false>(alpha, nullptr, i_start[r], i_end[r], weight, T, stop, T_last);
    if (lane == 0) { i_end[r] = stop; alphainv_last[r] = T_last; }
  }
}

VX_API int vx_alpha2weight(const float* alpha, const int64_t* ray_id, int64_t n_pts, int n_rays, float* weight, float* T,
                           float* alphainv_last, int64_t* i_start, int64_t* i_end, cudaStream_t st) {
  if (n_rays > 0) {
    k_a2w_init<<<vx_blocks(n_rays, 256), 256, 0, st>>>(n_rays, alphainv_last, i_start, i_end);
    int rc = vx_check_launch("vx_alpha2weight(init)");
    if (rc) return rc;
  }
  if (n_pts <= 0 || n_rays <= 0) return 0;  // :629
  k_a2w_segments<<<vx_blocks(n_pts, 256), 256, 0, st>>>(ray_id, n_pts, i_start, i_end);
  int rc = vx_check_launch("vx_alpha2weight(segments)");
  if (rc) return rc;
  const int blocks = min(vx_blocks((int64_t)n_rays * 32, 256), vx_num_sms() * 8);
  k_alpha2weight<<<blocks, 256, 0, st>>>(alpha, n_rays, weight, T, alphainv_last, i_start, i_end);
  return vx_check_launch("vx_alpha2weight");
}

// ---------------------------------------------------------------------------------------------
// alpha2weight backward (render_utils_kernel.cu:653-707): reverse recurrence
//   grad[i] = gw[i]*T[i] - back / (1 - alpha[i] + 1e-10);  back += gw[i]*w[i]
// `back` before step i is  back0 + sum_{j>i} gw[j]*w[j]  -- a suffix sum, evaluated here in the
// reference's order (sequential float FMAs from the ray's last sample backwards) by the same
// lock-step replay, so the gradient is bit-exact as well.
// ---------------------------------------------------------------------------------------------
template <bool kHasKeep>
__device__ __forceinline__ void a2w_ray_backward(const float* __restrict__ alpha, const float* __restrict__ weight,
                                                 const float* __restrict__ T, const uint8_t* __restrict__ keep,
                                                 int64_t i_s, int64_t i_e, float back0,
                                                 const float* __restrict__ grad_weights, float* __restrict__ grad) {
  const int lane = threadIdx.x & 31;
  float back_cum = back0;
  // chunks aligned to the segment END so the replay walks i_e-1, i_e-2, ...
  for (int64_t top = i_e; top > i_s; top -= 32) {
    const int64_t i = top - 1 - lane;  // lane 0 = last sample of the chunk
    const bool valid = i >= i_s;
    const bool k = kHasKeep ? (valid && keep[i]) : valid;
    const float gw = k ? grad_weights[i] : 0.f;
    const float w = k ? weight[i] : 0.f;
    const uint32_t kmask = __ballot_sync(0xffffffffu, k);
    float my_back = 0.f;
    const int cnt = (int)min((int64_t)32, top - i_s);
    for (int j = 0; j < cnt; ++j) {
      if (kHasKeep && !((kmask >> j) & 1u)) continue;
      if (lane == j) my_back = back_cum;
      const float gwj = __shfl_sync(0xffffffffu, gw, j);
      const float wj = __shfl_sync(0xffffffffu, w, j);
      back_cum += gwj * wj;
    }
    if (k) grad[i] = gw * T[i] - my_back / (1 - alpha[i] + 1e-10);
  }
}

__global__ void k_alpha2weight_backward(const float* __restrict__ alpha, const float* __restrict__ weight,
                                        const float* __restrict__ T, const float* __restrict__ alphainv_last,
                                        const int64_t* __restrict__ i_start, const int64_t* __restrict__ i_end,
                                        int n_rays, const float* __restrict__ grad_weights,
                                        const float* __restrict__ grad_last, float* __restrict__ grad) {
  const int warps_per_block = blockDim.x >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * warps_per_block) {
    const float back0 = grad_last[r] * alphainv_last[r];
    a2w_ray_backward<false>(alpha, weight, T, nullptr, i_start[r], i_end[r], back0, grad_weights, grad);
  }
}

VX_API int vx_alpha2weight_backward(const float* alpha, const float* weight, const float* T, const float* alphainv_last,
                                    const int64_t* i_start, const int64_t* i_end, int n_rays, int64_t n_pts,
                                    const float* grad_weights, const float* grad_last, float* grad, cudaStream_t st) {
  if (n_pts > 0) {
    cudaError_t e = cudaMemsetAsync(grad, 0, sizeof(float) * (size_t)n_pts, st);  // zeros_like (:684)
    if (e != cudaSuccess) { vx_set_error("vx_alpha2weight_backward", cudaGetErrorString(e)); return (int)e; }
  }
  if (n_rays <= 0 || n_pts <= 0) return 0;
  const int blocks = min(vx_blocks((int64_t)n_rays * 32, 256), vx_num_sms() * 8);
  k_alpha2weight_backward<<<blocks, 256, 0, st>>>(alpha, weight, T, alphainv_last, i_start, i_end, n_rays, grad_weights,
                                                  grad_last, grad);
  return vx_check_launch("vx_alpha2weight_backward");
}

// ---------------------------------------------------------------------------------------------
// Fused-path variants: int32 per-ray segment offsets (already known from the march scan), an
// optional per-sample keep flag (the alpha > fast_color_thres filter of voxurf_fine.py:647-654
// applied in place instead of compacting six tensors), and weight > thres flags + per-ray
// survivor counts for the second compaction (voxurf_fine.py:668-676).
// ---------------------------------------------------------------------------------------------
__global__ void k_a2w_seg(const float* __restrict__ alpha, const uint8_t* __restrict__ keep,
                          const int* __restrict__ seg_off, int n_rays, float w_thres, float* __restrict__ weight,
                          float* __restrict__ T, float* __restrict__ alphainv_last, int* __restrict__ i_end_out,
                          uint8_t* __restrict__ w_keep, int* __restrict__ w_count) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * warps_per_block) {
    const int64_t i_s = seg_off[r], i_e = seg_off[r + 1];
    int64_t stop;
    float T_last;
    if (keep) a2w_ray<true>(alpha, keep, i_s, i_e, weight, T, stop, T_last);
    else a2w_ray<false>(alpha, nullptr, i_s, i_e, weight, T, stop, T_last);
    __syncwarp();
    int cnt = 0;
    if (w_keep) {
      for (int64_t base = i_s; base < i_e; base += 32) {
        const int64_t i = base + lane;
        bool f = false;
        if (i < i_e) {
          f = weight[i] > w_thres;
          if (keep) f = f && keep[i];
          w_keep[i] = f;
        }
        cnt += __popc(__ballot_sync(0xffffffffu, f));
      }
    }
    if (lane == 0) {
      alphainv_last[r] = T_last;
      i_end_out[r] = (int)stop;
      if (w_count) w_count[r] = cnt;
    }
  }
}

VX_API int vx_alpha2weight_seg(const float* alpha, const uint8_t* keep, const int* seg_off, int n_rays, float w_thres,
                               float* weight, float* T, float* alphainv_last, int* i_end, uint8_t* w_keep, int* w_count,
                               cudaStream_t st) {
  if (n_rays <= 0) return 0;
  const int blocks = min(vx_blocks((int64_t)n_rays * 32, 256), vx_num_sms() * 8);
  k_a2w_seg<<<blocks, 256, 0, st>>>(alpha, keep, seg_off, n_rays, w_thres, weight, T, alphainv_last, i_end, w_keep, w_count);
  return vx_check_launch("vx_alpha2weight_seg");
}

__global__ void k_a2w_seg_backward(const float* __restrict__ alpha, const float* __restrict__ weight,
                                   const float* __restrict__ T, const uint8_t* __restrict__ keep,
                                   const float* __restrict__ alphainv_last, const int* __restrict__ seg_off,
                                   const int* __restrict__ i_end, int n_rays, const float* __restrict__ grad_weights,
                                   const float* __restrict__ grad_last, float* __restrict__ grad) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * warps_per_block) {
    const int64_t i_s = seg_off[r], i_seg_e = seg_off[r + 1], i_e = i_end[r];
    // samples past the early exit (and dropped ones) get zero gradient
    for (int64_t i = i_e + lane; i < i_seg_e; i += 32) grad[i] = 0.f;
    if (keep)
      for (int64_t i = i_s + lane; i < i_e; i += 32)
        if (!keep[i]) grad[i] = 0.f;
    const float back0 = grad_last[r] * alphainv_last[r];
    if (keep) a2w_ray_backward<true>(alpha, weight, T, keep, i_s, i_e, back0, grad_weights, grad);
    else a2w_ray_backward<false>(alpha, weight, T, nullptr, i_s, i_e, back0, grad_weights, grad);
  }
}

VX_API int vx_alpha2weight_seg_backward(const float* alpha, const float* weight, const float* T, const uint8_t* keep,
                                        const float* alphainv_last, const int* seg_off, const int* i_end, int n_rays,
                                        const float* grad_weights, const float* grad_last, float* grad, cudaStream_t st) {
  if (n_rays <= 0) return 0;
  const int blocks = min(vx_blocks((int64_t)n_rays * 32, 256), vx_num_sms() * 8);
  k_a2w_seg_backward<<<blocks, 256, 0, st>>>(alpha, weight, T, keep, alphainv_last, seg_off, i_end, n_rays, grad_weights,
                                             grad_last, grad);
  return vx_check_launch("vx_alpha2weight_seg_backward");
}

// ---------------------------------------------------------------------------------------------
// cumdist_thres (lib/cuda/ub360_utils_kernel.cu:13-47, the one native op of the unbounded "womask" models,
// lib/voxurf_womask_fine.py:867): per ray, walk the inter-sample distances, mark a sample when the running distance
// exceeds `thres`, and restart the running distance there.  The reference runs one THREAD per ray (strided, uncoalesced
// reads); here one WARP per ray reads 32 consecutive distances at a time and replays the recurrence in lock-step from
// shuffled values -- the same float additions in the same order, so the mask is bit-identical.
// ---------------------------------------------------------------------------------------------
__global__ void k_cumdist_thres(const float* __restrict__ dist, float thres, int n_rays, int n_pts, bool* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * warps_per_block) {
    const float* d = dist + (int64_t)r * n_pts;
    bool* m = mask + (int64_t)r * n_pts;
    float cum = 0.f;
    for (int base = 0; base < n_pts; base += 32) {
      const int i = base + lane;
      const float v = (i < n_pts) ? __ldg(d + i) : 0.f;
      bool mine = false;
      const int n = min(32, n_pts - base);
      for (int j = 0; j < n; ++j) {
        cum += __shfl_sync(0xffffffffu, v, j);
        const bool over = cum > thres;
        if (over) cum = 0.f;          // cum_dist *= float(!over)
        if (j == lane) mine = over;
      }
      if (i < n_pts) m[i] = mine;
    }
  }
}

VX_API int vx_cumdist_thres(const float* dist, float thres, int n_rays, int n_pts, bool* mask, cudaStream_t st) {
  if (n_rays <= 0 || n_pts <= 0) return 0;
  VX_REQUIRE(dist && mask, "vx_cumdist_thres", "null pointer");
  k_cumdist_thres<<<min(vx_blocks((int64_t)n_rays * 32, 256), vx_num_sms() * 16), 256, 0, st>>>(dist, thres, n_rays, n_pts, mask);
  return vx_check_launch("vx_cumdist_thres");
}
