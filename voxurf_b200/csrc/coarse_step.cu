// Fused, sync-free coarse-stage step (BASELINE config 2): the kernels that differ from the fine
// stage's (fused_step.cu) in one Voxurf.forward + loss + backward of lib/voxurf_coarse.py:513-619, run.py:604-639.
//   * one MLP: rgb_feat = [k0 C | xyz 3 | sin 3P | cos 3P | view 3 | sin 3V | cos 3V | normal 3]   (lib/voxurf_coarse.py:552-571)
//     with normal = gradient / (|gradient| + 1e-5), gradient = trilinear sample of the FD gradient GRID of the raw sdf
//   * compositing with the accumulated weight: rgb_marched = sum w rgb + (1 - sum w) bg, clamped       (:575-583)
// The M2-level pass (grid gathers, NeuS alpha, the two alpha2weight passes -- the second on the survivors of the
// weight > thres filter, hazard 8) runs on the operator-level entry points with device-side counts; what is here are the
// row-level kernels.  Every count lives on the device: the step is CUDA-graph capturable like the fine one.
#include "common.cuh"
#include "taps.cuh"

struct VxCoarseLayout {
  int P, V, C, ld;
};

// ---------------------------------------------------------------------------------------------
// MLP input rows.  Four threads per row: the sin / cos pairs are dealt round-robin, sub-thread 1 gathers k0,
// sub-thread 2 writes the normal.  Rows >= M4 (up to capacity) are zero-filled.
// ---------------------------------------------------------------------------------------------
template <int kC>
__global__ void k_coarse_row_features(VxGrid gk, const float* __restrict__ k0_grid, VxPts pts, const int* __restrict__ idx4,
                                      const int* __restrict__ n_rows_dev, int capacity, const float* __restrict__ viewdirs,
                                      const float* __restrict__ grad_s, VxCoarseLayout lay, float* __restrict__ X) {
  const int n = min(*n_rows_dev, capacity);
  for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < capacity * 4; item += gridDim.x * blockDim.x) {
    const int row = item >> 2, sub = item & 3;
    float* x = X + (int64_t)row * lay.ld;
    if (row >= n) {
      for (int c = sub; c < lay.ld; c += 4) x[c] = 0.f;
      continue;
    }
    const int i = idx4[row];
    float p[3];
    vx_load_pt(pts, i, p[0], p[1], p[2]);
    const int r = pts.ray_id[i];
    float xn[3], vd[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      xn[d] = __fdiv_rn(__fsub_rn(p[d], gk.min[d]), __fsub_rn(gk.max[d], gk.min[d]));
      vd[d] = viewdirs[3 * r + d];
    }
    int c = kC;
    if (sub == 0) {
#pragma unroll
      for (int d = 0; d < 3; ++d) x[c + d] = xn[d];
    }
    c += 3;
    for (int q = sub; q < 3 * lay.P; q += 4) {
      const int d = q / lay.P, f = q - d * lay.P;
      float sn, cs;
      sincosf(__fmul_rn(d == 0 ? xn[0] : (d == 1 ? xn[1] : xn[2]), (float)(1 << f)), &sn, &cs);
      x[c + q] = sn; x[c + 3 * lay.P + q] = cs;
    }
    c += 6 * lay.P;
    if (sub == 3) {
#pragma unroll
      for (int d = 0; d < 3; ++d) x[c + d] = vd[d];
    }
    c += 3;
    for (int q = sub; q < 3 * lay.V; q += 4) {
      const int d = q / lay.V, f = q - d * lay.V;
      float sn, cs;
      sincosf(__fmul_rn(d == 0 ? vd[0] : (d == 1 ? vd[1] : vd[2]), (float)(1 << f)), &sn, &cs);
      x[c + q] = sn; x[c + 3 * lay.V + q] = cs;
    }
    c += 6 * lay.V;
    if (sub == 2) {
      const float gx = grad_s[3 * i], gy = grad_s[3 * i + 1], gz = grad_s[3 * i + 2];
      const float den = sqrtf(gx * gx + gy * gy + gz * gz) + 1e-5f;     // lib/voxurf_coarse.py:568
      x[c] = gx / den; x[c + 1] = gy / den; x[c + 2] = gz / den;
      for (int cc = c + 3; cc < lay.ld; ++cc) x[cc] = 0.f;
    }
    if (sub == 1) {    // k0 trilinear gather (DenseGrid.forward lib/grid.py:47-58)
      float ix, iy, iz;
      point_to_index(gk, p[0], p[1], p[2], ix, iy, iz);
      VxTap t;
      vx_make_tap(ix, iy, iz, gk.X, gk.Y, gk.Z, t);
      const int64_t V = (int64_t)gk.X * gk.Y * gk.Z;
      float acc[kC];
#pragma unroll
      for (int cc = 0; cc < kC; ++cc) acc[cc] = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (t.off[k] < 0) continue;
        if (gk.cl) {
          const float* src = k0_grid + (int64_t)t.off[k] * kC;
#pragma unroll
          for (int cc = 0; cc < kC; cc += 2) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(src + cc));
            acc[cc] += v.x * t.w[k]; acc[cc + 1] += v.y * t.w[k];
          }
        } else {
#pragma unroll
          for (int cc = 0; cc < kC; ++cc) acc[cc] += __ldg(k0_grid + cc * V + t.off[k]) * t.w[k];
        }
      }
#pragma unroll
      for (int cc = 0; cc < kC; ++cc) x[cc] = acc[cc];
    }
  }
}

VX_API int vx_coarse_row_features(const float* k0_grid, int X, int Y, int Z, int C, int k0_channels_last,
                                  const float* xyz_min_host, const float* xyz_max_host, const int* ray_id, const int* step_id,
                                  const float* rays_start, const float* rays_dir, float stepdist, const int* idx4,
                                  const int* n_rows_dev, int capacity, const float* viewdirs, const float* grad_s, int P, int V,
                                  int ld, float* Xrows, cudaStream_t st) {
  if (capacity <= 0) return 0;
  VX_REQUIRE(C == 6 || C == 12, "vx_coarse_row_features", "k0 channels must be 6 or 12");
  VX_REQUIRE(ld >= C + 3 + 6 * P + 3 + 6 * V + 3 && P >= 0 && V >= 0, "vx_coarse_row_features", "bad layout");
  const VxGrid gk = make_grid(X, Y, Z, C, k0_channels_last, xyz_min_host, xyz_max_host);
  const VxPts pts{nullptr, ray_id, step_id, rays_start, rays_dir, stepdist};
  const VxCoarseLayout lay{P, V, C, ld};
  const int blocks = min(vx_blocks((int64_t)capacity * 4, 128), vx_num_sms() * 32);
  if (C == 6) k_coarse_row_features<6><<<blocks, 128, 0, st>>>(gk, k0_grid, pts, idx4, n_rows_dev, capacity, viewdirs, grad_s, lay, Xrows);
  else k_coarse_row_features<12><<<blocks, 128, 0, st>>>(gk, k0_grid, pts, idx4, n_rows_dev, capacity, viewdirs, grad_s, lay, Xrows);
  return vx_check_launch("vx_coarse_row_features");
}

// backward of the rows: k0 scatter (each of a row's four threads takes two corners) and the normal's adjoint added to
// d_grad_s of the row's sample (every row owns a distinct sample: plain read-modify-write)
template <int kC>
__global__ void k_coarse_row_backward(VxGrid gk, VxPts pts, const int* __restrict__ idx4, const int* __restrict__ n_rows_dev,
                                      int capacity, const float* __restrict__ grad_s, VxCoarseLayout lay,
                                      const float* __restrict__ dX, float* __restrict__ d_grad_s, float* __restrict__ k0_grad) {
  const int n = min(*n_rows_dev, capacity);
  const int col_n = kC + 3 + 6 * lay.P + 3 + 6 * lay.V;
  for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < n * 4; item += gridDim.x * blockDim.x) {
    const int row = item >> 2, sub = item & 3;
    const float* g = dX + (int64_t)row * lay.ld;
    const int i = idx4[row];
    if (sub == 3) {
      const float dn[3] = {g[col_n], g[col_n + 1], g[col_n + 2]};
      const float gr[3] = {grad_s[3 * i], grad_s[3 * i + 1], grad_s[3 * i + 2]};
      // y = g / (|g| + eps):  dg = dy / (n + eps) - g <dy, g> / ((n + eps)^2 n)     (zero second term at n == 0)
      const float nrm = sqrtf(gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2]);
      const float den = nrm + 1e-5f;
      const float dot = dn[0] * gr[0] + dn[1] * gr[1] + dn[2] * gr[2];
      const float k = (nrm > 0.f) ? dot / (den * den * nrm) : 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a) d_grad_s[3 * i + a] += dn[a] / den - gr[a] * k;
    }
    float go[kC];
    bool any = false;
#pragma unroll
    for (int c = 0; c < kC; ++c) { go[c] = g[c]; any |= go[c] != 0.f; }
    if (!any) continue;
    float p[3], ix, iy, iz;
    vx_load_pt(pts, i, p[0], p[1], p[2]);
    point_to_index(gk, p[0], p[1], p[2], ix, iy, iz);
    VxTap t;
    vx_make_tap(ix, iy, iz, gk.X, gk.Y, gk.Z, t);
    const int64_t V = (int64_t)gk.X * gk.Y * gk.Z;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if ((k >> 1) != sub || t.off[k] < 0) continue;
      if (gk.cl) {
        float* dst = k0_grad + (int64_t)t.off[k] * kC;
#pragma unroll
        for (int c = 0; c < kC; c += 2) atomicAdd(reinterpret_cast<float2*>(dst + c), make_float2(go[c] * t.w[k], go[c + 1] * t.w[k]));
      } else {
#pragma unroll
        for (int c = 0; c < kC; ++c) atomicAdd(k0_grad + c * V + t.off[k], go[c] * t.w[k]);
      }
    }
  }
}

VX_API int vx_coarse_row_backward(int X, int Y, int Z, int C, int k0_channels_last, const float* xyz_min_host,
                                  const float* xyz_max_host, const int* ray_id, const int* step_id, const float* rays_start,
                                  const float* rays_dir, float stepdist, const int* idx4, const int* n_rows_dev, int capacity,
                                  const float* grad_s, int P, int V, int ld, const float* dX, float* d_grad_s, float* k0_grad,
                                  cudaStream_t st) {
  if (capacity <= 0) return 0;
  VX_REQUIRE(C == 6 || C == 12, "vx_coarse_row_backward", "k0 channels must be 6 or 12");
  const VxGrid gk = make_grid(X, Y, Z, C, k0_channels_last, xyz_min_host, xyz_max_host);
  const VxPts pts{nullptr, ray_id, step_id, rays_start, rays_dir, stepdist};
  const VxCoarseLayout lay{P, V, C, ld};
  const int blocks = vx_num_sms() * 16;
  if (C == 6) k_coarse_row_backward<6><<<blocks, 128, 0, st>>>(gk, pts, idx4, n_rows_dev, capacity, grad_s, lay, dX, d_grad_s, k0_grad);
  else k_coarse_row_backward<12><<<blocks, 128, 0, st>>>(gk, pts, idx4, n_rows_dev, capacity, grad_s, lay, dX, d_grad_s, k0_grad);
  return vx_check_launch("vx_coarse_row_backward");
}

// ---------------------------------------------------------------------------------------------
// Compositing + losses + their backward (lib/voxurf_coarse.py:573-583, run.py:604-610): one warp per ray, deterministic
// segmented sums.  rgb_marched = clamp(sum_i w_i sigmoid(logit_i) + (1 - sum_i w_i) bg, 0, 1); loss = w_main * mse
// (+ the entropy term on the LAST ray's alphainv_last, run.py:608).  train: d_logit (rows), d_w (scattered to the
// sample of each row; samples that are not rows keep the zero the caller put there), d_last, per-ray loss.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float c_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void k_coarse_composite_loss(const float* __restrict__ logit, int ld_out, const int* __restrict__ idx4,
                                        const int* __restrict__ off4, int capacity, const float* __restrict__ weight_s,
                                        const float* __restrict__ alphainv_last, const float* __restrict__ target, int n_rays,
                                        float w_main, float w_ent, float ent_scale, float bg, int train,
                                        float* __restrict__ rgb_marched, float* __restrict__ d_logit, float* __restrict__ d_w_s,
                                        float* __restrict__ d_last, float* __restrict__ loss_ray) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const float inv_3n = 1.f / (3.f * (float)n_rays);
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * warps_per_block) {
    const int b = min(off4[r], capacity), e = min(off4[r + 1], capacity);
    float s[3] = {0.f, 0.f, 0.f}, sw = 0.f;
    for (int row = b + lane; row < e; row += 32) {
      const float w = weight_s[idx4[row]];
      sw += w;
#pragma unroll
      for (int c = 0; c < 3; ++c) s[c] += w * sigmoidf_(logit[(int64_t)row * ld_out + c]);
    }
    sw = c_warp_sum(sw);
    float g[3], lsum = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float raw = c_warp_sum(s[c]) + (1.f - sw) * bg;
      const float m = fminf(fmaxf(raw, 0.f), 1.f);
      if (lane == 0) rgb_marched[3 * r + c] = m;
      const float tg = target ? target[3 * r + c] : 0.f;
      const float err = m - tg;
      g[c] = (raw >= 0.f && raw <= 1.f) ? w_main * 2.f * err * inv_3n : 0.f;
      lsum += w_main * err * err * inv_3n;
    }
    if (!train) continue;
    float dl = 0.f;
    if (r == n_rays - 1 && w_ent > 0.f && ent_scale != 0.f) {
      const float al = alphainv_last[r];
      const float pc = fminf(fmaxf(al, 1e-6f), 1.f - 1e-6f);
      lsum += ent_scale * w_ent * (-(pc * logf(pc) + (1.f - pc) * logf(1.f - pc)));
      if (al >= 1e-6f && al <= 1.f - 1e-6f) dl = ent_scale * w_ent * (logf(1.f - pc) - logf(pc));
    }
    if (lane == 0) { d_last[r] = dl; loss_ray[r] = lsum; }
    for (int row = b + lane; row < e; row += 32) {
      const int i = idx4[row];
      const float w = weight_s[i];
      float dw = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float rgb = sigmoidf_(logit[(int64_t)row * ld_out + c]);
        dw += g[c] * (rgb - bg);
        d_logit[(int64_t)row * ld_out + c] = w * g[c] * rgb * (1.f - rgb);
      }
      d_w_s[i] = dw;
    }
  }
}

__global__ void k_coarse_zero_tail(float* __restrict__ a, int ld, const int* __restrict__ n_rows_dev, int capacity) {
  const int n = min(*n_rows_dev, capacity);
  for (int64_t t = (int64_t)n * ld + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (int64_t)capacity * ld;
       t += (int64_t)gridDim.x * blockDim.x)
    a[t] = 0.f;
}

VX_API int vx_coarse_composite_loss(const float* logit, int ld_out, const int* idx4, const int* off4, int capacity,
                                    const float* weight_s, const float* alphainv_last, const float* target, int n_rays,
                                    float w_main, float w_ent, float ent_scale, float bg, int train, float* rgb_marched,
                                    float* d_logit, float* d_w_s, float* d_last, float* loss_ray, cudaStream_t st) {
  if (n_rays <= 0) return 0;
  if (train) {
    k_coarse_zero_tail<<<vx_num_sms() * 2, 256, 0, st>>>(d_logit, ld_out, off4 + n_rays, capacity);
    int rc = vx_check_launch("vx_coarse_composite_loss(tail)");
    if (rc) return rc;
  }
  const int blocks = min(vx_blocks((int64_t)n_rays * 32, 256), vx_num_sms() * 8);
  k_coarse_composite_loss<<<blocks, 256, 0, st>>>(logit, ld_out, idx4, off4, capacity, weight_s, alphainv_last, target, n_rays,
                                                  w_main, w_ent, ent_scale, bg, train, rgb_marched, d_logit, d_w_s, d_last, loss_ray);
  return vx_check_launch("vx_coarse_composite_loss");
}

// y += a * x over n floats (the autograd-form regularisers return their gradient; the step adds it, scaled)
__global__ void k_axpy(const float* __restrict__ x, float a, int64_t n, float* __restrict__ y) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) y[t] += a * x[t];
}

VX_API int vx_axpy(const float* x, float a, int64_t n, float* y, cudaStream_t st) {
  if (n <= 0) return 0;
  k_axpy<<<(int)min((int64_t)vx_blocks(n, 256), (int64_t)vx_num_sms() * 16), 256, 0, st>>>(x, a, n, y);
  return vx_check_launch("vx_axpy");
}
