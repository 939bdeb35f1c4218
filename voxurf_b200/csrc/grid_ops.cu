// Trilinear grid gathers and their backward scatters.
//
// Reference behaviour being matched (all via F.grid_sample, 'bilinear', align_corners=True, zeros
// padding): DenseGrid.forward lib/grid.py:47-58; Voxurf.grid_sampler lib/voxurf_fine.py:502-534;
// Voxurf.sample_sdfs lib/voxurf_fine.py:537-577; coarse grid_sampler lib/voxurf_coarse.py:435-452.
// The reference issues one grid_sample per tap (7 at M2, 24 + C at M4) plus ~15 elementwise
// kernels; here one thread evaluates every tap of a sample and writes the final features.
// Backward: the reference relies on ATen's grid_sampler_3d_backward (atomicAdd per corner per tap,
// into a freshly zero-allocated grad tensor).  Here gradients are first combined per sample in
// registers, exact zeros are dropped (samples behind the early-termination point have exactly zero
// gradient), and what is left goes out as fp32 atomics into a persistent, already-zero grad grid.
#include "common.cuh"
#include "taps.cuh"

// ---------------------------------------------------------------------------------------------
// generic C-channel gather:  out[p, c] = trilinear(grid[c], xyz[p])
// ---------------------------------------------------------------------------------------------
template <int kC>
__global__ void k_grid_gather(VxGrid g, const float* __restrict__ grid, VxPts pts, const int* __restrict__ n_dev,
                              int64_t n_host, float* __restrict__ out) {
  const int64_t n = vx_count(n_dev, n_host);
  const int C = kC > 0 ? kC : g.C;
  const int64_t V = (int64_t)g.X * g.Y * g.Z;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    float px, py, pz, ix, iy, iz;
    vx_load_pt(pts, p, px, py, pz);
    point_to_index(g, px, py, pz, ix, iy, iz);
    VxTap t;
    vx_make_tap(ix, iy, iz, g.X, g.Y, g.Z, t);
    if (g.cl) {
      float acc[kC > 0 ? kC : 1];
      if (kC > 0) {
#pragma unroll
        for (int c = 0; c < kC; ++c) acc[c] = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (t.off[k] < 0) continue;
          const float* src = grid + (int64_t)t.off[k] * kC;
          if (kC % 4 == 0) {
#pragma unroll
            for (int c = 0; c < kC; c += 4) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(src + c));
              acc[c] += v.x * t.w[k]; acc[c + 1] += v.y * t.w[k]; acc[c + 2] += v.z * t.w[k]; acc[c + 3] += v.w * t.w[k];
            }
          } else if (kC % 2 == 0) {
#pragma unroll
            for (int c = 0; c < kC; c += 2) {
              const float2 v = __ldg(reinterpret_cast<const float2*>(src + c));
              acc[c] += v.x * t.w[k]; acc[c + 1] += v.y * t.w[k];
            }
          } else {
#pragma unroll
            for (int c = 0; c < kC; ++c) acc[c] += __ldg(src + c) * t.w[k];
          }
        }
#pragma unroll
        for (int c = 0; c < kC; ++c) out[p * kC + c] = acc[c];
      } else {
        for (int c = 0; c < C; ++c) {
          float a = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (t.off[k] >= 0) a += __ldg(grid + (int64_t)t.off[k] * C + c) * t.w[k];
          out[p * C + c] = a;
        }
      }
    } else {
      for (int c = 0; c < C; ++c) out[p * C + c] = vx_tap_eval(grid + c * V, t);
    }
  }
}

template <int kC>
__global__ void k_grid_gather_bwd(VxGrid g, VxPts pts, const int* __restrict__ n_dev, int64_t n_host,
                                  const float* __restrict__ grad_out, float* __restrict__ grad_grid,
                                  uint32_t* __restrict__ touched) {
  const int64_t n = vx_count(n_dev, n_host);
  const int C = kC > 0 ? kC : g.C;
  const int64_t V = (int64_t)g.X * g.Y * g.Z;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    float go[kC > 0 ? kC : 1];
    bool any = false;
    if (kC > 0) {
#pragma unroll
      for (int c = 0; c < kC; ++c) { go[c] = grad_out[p * kC + c]; any |= (go[c] != 0.f); }
      if (!any) continue;
    }
    float px, py, pz, ix, iy, iz;
    vx_load_pt(pts, p, px, py, pz);
    point_to_index(g, px, py, pz, ix, iy, iz);
    VxTap t;
    vx_make_tap(ix, iy, iz, g.X, g.Y, g.Z, t);
    if (touched) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (t.off[k] >= 0) atomicOr(touched + (t.off[k] >> 5), 1u << (t.off[k] & 31));
    }
    if (g.cl && kC > 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (t.off[k] < 0) continue;
        float* dst = grad_grid + (int64_t)t.off[k] * kC;
        if (kC % 4 == 0) {
#pragma unroll
          for (int c = 0; c < kC; c += 4)
            atomicAdd(reinterpret_cast<float4*>(dst + c),
                      make_float4(go[c] * t.w[k], go[c + 1] * t.w[k], go[c + 2] * t.w[k], go[c + 3] * t.w[k]));
        } else if (kC % 2 == 0) {
#pragma unroll
          for (int c = 0; c < kC; c += 2)
            atomicAdd(reinterpret_cast<float2*>(dst + c), make_float2(go[c] * t.w[k], go[c + 1] * t.w[k]));
        } else {
#pragma unroll
          for (int c = 0; c < kC; ++c) atomicAdd(dst + c, go[c] * t.w[k]);
        }
      }
    } else if (g.cl) {
      for (int c = 0; c < C; ++c) {
        const float gv = grad_out[p * C + c];
        if (gv == 0.f) continue;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (t.off[k] >= 0) atomicAdd(grad_grid + (int64_t)t.off[k] * C + c, gv * t.w[k]);
      }
    } else {
      for (int c = 0; c < C; ++c) vx_tap_scatter(grad_grid + c * V, t, kC > 0 ? go[c] : grad_out[p * C + c]);
    }
  }
}

#define VX_DISPATCH_C(C, cl, CALL)                  \
  if ((cl) && (C) == 12) { CALL(12); }              \
  else if ((cl) && (C) == 6) { CALL(6); }           \
  else if ((cl) && (C) == 3) { CALL(3); }           \
  else if ((cl) && (C) == 4) { CALL(4); }           \
  else { CALL(0); }

// channels-last scatter with four threads per point (corners 2 sub, 2 sub + 1 each): the 24 vector atomics of a 12-channel
// point are spread over four lanes instead of serialised in one (the multi-GPU step re-scatters world x rows with this)
template <int kC>
__global__ void k_grid_gather_bwd_cl4(VxGrid g, VxPts pts, const int* __restrict__ n_dev, int64_t n_host,
                                      const float* __restrict__ grad_out, float* __restrict__ grad_grid,
                                      uint32_t* __restrict__ touched) {
  const int64_t n = vx_count(n_dev, n_host);
  for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < n * 4; item += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = item >> 2;
    const int sub = (int)(item & 3);
    float go[kC];
    bool any = false;
#pragma unroll
    for (int c = 0; c < kC; ++c) { go[c] = grad_out[p * kC + c]; any |= (go[c] != 0.f); }
    if (!any) continue;
    float px, py, pz, ix, iy, iz;
    vx_load_pt(pts, p, px, py, pz);
    point_to_index(g, px, py, pz, ix, iy, iz);
    VxTap t;
    vx_make_tap(ix, iy, iz, g.X, g.Y, g.Z, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if ((k >> 1) != sub || t.off[k] < 0) continue;
      if (touched) atomicOr(touched + (t.off[k] >> 5), 1u << (t.off[k] & 31));
      float* dst = grad_grid + (int64_t)t.off[k] * kC;
      if (kC % 4 == 0) {
#pragma unroll
        for (int c = 0; c < kC; c += 4)
          atomicAdd(reinterpret_cast<float4*>(dst + c),
                    make_float4(go[c] * t.w[k], go[c + 1] * t.w[k], go[c + 2] * t.w[k], go[c + 3] * t.w[k]));
      } else {
#pragma unroll
        for (int c = 0; c < kC; c += 2)
          atomicAdd(reinterpret_cast<float2*>(dst + c), make_float2(go[c] * t.w[k], go[c + 1] * t.w[k]));
      }
    }
  }
}

// Data-parallel k0 exchange, owner side: the all-gathered rows of EVERY rank -- recv[r] = [cap x 3 positions |
// cap x C gradient rows | row count (int32) + 3 pad] -- scattered in one launch, restricted to the corners whose voxel lies
// in the linear voxel range [v_lo, v_hi) (the X-slab this rank owns; the whole grid when the k0 grid is not sharded).
// Four threads per row like k_grid_gather_bwd_cl4.
template <int kC>
__global__ void k_k0_rows_scatter(VxGrid g, const float* __restrict__ recv, int world, int cap, int v_lo, int v_hi,
                                  float* __restrict__ grad_grid, uint32_t* __restrict__ touched) {
  const int64_t stride_r = (int64_t)cap * (3 + kC) + 4;
  const int64_t n_items = (int64_t)world * cap * 4;
  for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(item / ((int64_t)cap * 4));
    const int64_t q = item - (int64_t)r * cap * 4;
    const int p = (int)(q >> 2), sub = (int)(q & 3);
    const float* base = recv + r * stride_r;
    const int n = min(__ldg(reinterpret_cast<const int*>(base + (int64_t)cap * (3 + kC))), cap);
    if (p >= n) continue;
    const float* gp = base + (int64_t)cap * 3 + (int64_t)p * kC;
    float go[kC];
    bool any = false;
#pragma unroll
    for (int c = 0; c < kC; ++c) { go[c] = gp[c]; any |= (go[c] != 0.f); }
    if (!any) continue;
    float ix, iy, iz;
    point_to_index(g, base[3 * p], base[3 * p + 1], base[3 * p + 2], ix, iy, iz);
    VxTap t;
    vx_make_tap(ix, iy, iz, g.X, g.Y, g.Z, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if ((k >> 1) != sub || t.off[k] < v_lo || t.off[k] >= v_hi) continue;
      if (touched) atomicOr(touched + (t.off[k] >> 5), 1u << (t.off[k] & 31));
      float* dst = grad_grid + (int64_t)t.off[k] * kC;
      if (kC % 4 == 0) {
#pragma unroll
        for (int c = 0; c < kC; c += 4)
          atomicAdd(reinterpret_cast<float4*>(dst + c),
                    make_float4(go[c] * t.w[k], go[c + 1] * t.w[k], go[c + 2] * t.w[k], go[c + 3] * t.w[k]));
      } else {
#pragma unroll
        for (int c = 0; c < kC; c += 2)
          atomicAdd(reinterpret_cast<float2*>(dst + c), make_float2(go[c] * t.w[k], go[c + 1] * t.w[k]));
      }
    }
  }
}

VX_API int vx_k0_rows_scatter(int X, int Y, int Z, int C, const float* xyz_min_host, const float* xyz_max_host,
                              const float* recv, int world, int cap, int x_lo, int x_hi, float* grad_grid,
                              uint32_t* touched, cudaStream_t st) {
  if (world <= 0 || cap <= 0) return 0;
  VX_REQUIRE(C == 12 || C == 6, "vx_k0_rows_scatter", "channels-last k0 grid with 6 or 12 channels");
  VX_REQUIRE((int64_t)X * Y * Z < ((int64_t)1 << 31) && 0 <= x_lo && x_lo <= x_hi && x_hi <= X, "vx_k0_rows_scatter", "bad grid / slab");
  const VxGrid g = make_grid(X, Y, Z, C, 1, xyz_min_host, xyz_max_host);
  const int blocks = (int)min((int64_t)vx_blocks((int64_t)world * cap * 4, 256), (int64_t)vx_num_sms() * 16);
  if (C == 12) k_k0_rows_scatter<12><<<blocks, 256, 0, st>>>(g, recv, world, cap, x_lo * Y * Z, x_hi * Y * Z, grad_grid, touched);
  else k_k0_rows_scatter<6><<<blocks, 256, 0, st>>>(g, recv, world, cap, x_lo * Y * Z, x_hi * Y * Z, grad_grid, touched);
  return vx_check_launch("vx_k0_rows_scatter");
}

// xyz_min / xyz_max are HOST float[3] (grid geometry is static model configuration).
VX_API int vx_grid_gather(const float* grid, int X, int Y, int Z, int C, int channels_last, const float* xyz_min_host,
                          const float* xyz_max_host, const float* xyz, const int* ray_id, const int* step_id,
                          const float* rays_start, const float* rays_dir, float stepdist, const int* n_dev,
                          int64_t n_host, float* out, cudaStream_t st) {
  if (!n_dev && n_host <= 0) return 0;
  VX_REQUIRE((int64_t)X * Y * Z * (int64_t)C < ((int64_t)1 << 40), "vx_grid_gather", "grid too large");
  VX_REQUIRE((int64_t)X * Y * Z < ((int64_t)1 << 31), "vx_grid_gather", "more than 2^31 voxels");
  const VxGrid g = make_grid(X, Y, Z, C, channels_last, xyz_min_host, xyz_max_host);
  const VxPts pts{xyz, ray_id, step_id, rays_start, rays_dir, stepdist};
  const int blocks = launch_blocks(n_dev, n_host);
#define CALL(KC) k_grid_gather<KC><<<blocks, 256, 0, st>>>(g, grid, pts, n_dev, n_host, out)
  VX_DISPATCH_C(C, channels_last, CALL)
#undef CALL
  return vx_check_launch("vx_grid_gather");
}

VX_API int vx_grid_gather_backward(int X, int Y, int Z, int C, int channels_last, const float* xyz_min_host,
                                   const float* xyz_max_host, const float* xyz, const int* ray_id, const int* step_id,
                                   const float* rays_start, const float* rays_dir, float stepdist, const int* n_dev,
                                   int64_t n_host, const float* grad_out, float* grad_grid, uint32_t* touched,
                                   cudaStream_t st) {
  if (!n_dev && n_host <= 0) return 0;
  const VxGrid g = make_grid(X, Y, Z, C, channels_last, xyz_min_host, xyz_max_host);
  const VxPts pts{xyz, ray_id, step_id, rays_start, rays_dir, stepdist};
  if (channels_last && (C == 12 || C == 6)) {
    const int blocks4 = n_dev ? vx_num_sms() * 8 : (int)min((int64_t)vx_blocks(n_host * 4, 256), (int64_t)vx_num_sms() * 16);
    if (C == 12) k_grid_gather_bwd_cl4<12><<<blocks4, 256, 0, st>>>(g, pts, n_dev, n_host, grad_out, grad_grid, touched);
    else k_grid_gather_bwd_cl4<6><<<blocks4, 256, 0, st>>>(g, pts, n_dev, n_host, grad_out, grad_grid, touched);
    return vx_check_launch("vx_grid_gather_backward");
  }
  const int blocks = launch_blocks(n_dev, n_host);
#define CALL(KC) k_grid_gather_bwd<KC><<<blocks, 256, 0, st>>>(g, pts, n_dev, n_host, grad_out, grad_grid, touched)
  VX_DISPATCH_C(C, channels_last, CALL)
#undef CALL
  return vx_check_launch("vx_grid_gather_backward");
}

// ---------------------------------------------------------------------------------------------
// SDF taps: centre value + L displaced finite differences (lib/voxurf_fine.py:537-577).
//   ind      = unnormalised index of the point (z,y,x order in the reference, :547)
//   tap(a,s,l) at ind + s*d_l along axis a, clamped to [0, size-1] (:552-556), then pushed through the
//   reference's normalise -> grid_sample unnormalise round trip (:558-559) so coordinates agree to the ulp
//   grad(a,l) = (f+ - f-) / (clamped index distance) / voxel_size (:562-566), optional per-l L2 normalise
// Output order: `xyz_order=0` reference sample_sdfs layout, feat[(a*2+s)*L + l], grad[a*L + l], a = z,y,x;
// `xyz_order=1` (only L == 1) Voxurf.grid_sampler layout, feat = x-,x+,y-,y+,z-,z+, grad = gx,gy,gz (:525-526).
// ---------------------------------------------------------------------------------------------
__global__ void k_sdf_taps(VxGrid g, const float* __restrict__ grid, VxPts pts, const int* __restrict__ n_dev,
                           int64_t n_host, VxDisp disp, float voxel_size, int use_grad_norm, int xyz_order,
                           float* __restrict__ out_sdf, float* __restrict__ out_feat, float* __restrict__ out_grad) {
  const int64_t n = vx_count(n_dev, n_host);
  const int L = disp.L;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    float px, py, pz;
    vx_load_pt(pts, p, px, py, pz);
    VxTap t;
    if (out_sdf) {  // plain trilinear at the point (voxurf_fine.py:516-519): no clamp, zeros padding
      float ix, iy, iz;
      point_to_index(g, px, py, pz, ix, iy, iz);
      vx_make_tap(ix, iy, iz, g.X, g.Y, g.Z, t);
      out_sdf[p] = vx_tap_eval(grid, t);
    }
    if (L == 0) continue;
    SdfTapCoords s;
    sdf_tap_setup(g, px, py, pz, s);
    for (int l = 0; l < L; ++l) {
      float gr[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        float ix, iy, iz;
        const float cm = sdf_tap_coords(g, s, a, -disp.d[l], ix, iy, iz);
        vx_make_tap(ix, iy, iz, g.X, g.Y, g.Z, t);
        const float fm = vx_tap_eval(grid, t);
        const float cp = sdf_tap_coords(g, s, a, disp.d[l], ix, iy, iz);
        vx_make_tap(ix, iy, iz, g.X, g.Y, g.Z, t);
        const float fp = vx_tap_eval(grid, t);
        gr[a] = __fdiv_rn(__fdiv_rn(__fsub_rn(fp, fm), __fsub_rn(cp, cm)), voxel_size);
        if (out_feat) {
          const int aa = xyz_order ? (2 - a) : a;
          out_feat[p * 6 * L + (aa * 2 + 0) * L + l] = fm;
          out_feat[p * 6 * L + (aa * 2 + 1) * L + l] = fp;
        }
      }
      if (use_grad_norm) {
        const float nrm = sqrtf(gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2]) + 1e-5f;
        gr[0] = gr[0] / nrm; gr[1] = gr[1] / nrm; gr[2] = gr[2] / nrm;
      }
      if (out_grad) {
#pragma unroll
        for (int a = 0; a < 3; ++a) out_grad[p * 3 * L + (xyz_order ? (2 - a) : a) * L + l] = gr[a];
      }
    }
  }
}

// backward of k_sdf_taps.  grad_* may be nullptr.  For use_grad_norm the raw gradient is recomputed.
__global__ void k_sdf_taps_bwd(VxGrid g, const float* __restrict__ grid, VxPts pts, const int* __restrict__ n_dev,
                               int64_t n_host, VxDisp disp, float voxel_size, int use_grad_norm, int xyz_order,
                               const float* __restrict__ grad_sdf, const float* __restrict__ grad_feat,
                               const float* __restrict__ grad_grad, float* __restrict__ grad_grid) {
  const int64_t n = vx_count(n_dev, n_host);
  const int L = disp.L;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    // cheap early-out: a sample whose incoming gradients are all exactly zero contributes nothing
    bool any = grad_sdf && grad_sdf[p] != 0.f;
    if (!any && grad_feat)
      for (int i = 0; i < 6 * L && !any; ++i) any = grad_feat[p * 6 * L + i] != 0.f;
    if (!any && grad_grad)
      for (int i = 0; i < 3 * L && !any; ++i) any = grad_grad[p * 3 * L + i] != 0.f;
    if (!any) continue;
    float px, py, pz;
    vx_load_pt(pts, p, px, py, pz);
    VxTap t;
    if (grad_sdf && grad_sdf[p] != 0.f) {
      float ix, iy, iz;
      point_to_index(g, px, py, pz, ix, iy, iz);
      vx_make_tap(ix, iy, iz, g.X, g.Y, g.Z, t);
      vx_tap_scatter(grad_grid, t, grad_sdf[p]);
    }
    if (L == 0) continue;
    SdfTapCoords s;
    sdf_tap_setup(g, px, py, pz, s);
    for (int l = 0; l < L; ++l) {
      float dgr[3] = {0.f, 0.f, 0.f};
      if (grad_grad) {
#pragma unroll
        for (int a = 0; a < 3; ++a) dgr[a] = grad_grad[p * 3 * L + (xyz_order ? (2 - a) : a) * L + l];
        if (use_grad_norm && (dgr[0] != 0.f || dgr[1] != 0.f || dgr[2] != 0.f)) {
          // y = g / (|g| + eps):  dg = dy/(n+eps) - g * <dy,g> / ((n+eps)^2 * n)   (zero second term at n == 0)
          float gr[3];
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            float ix, iy, iz;
            const float cm = sdf_tap_coords(g, s, a, -disp.d[l], ix, iy, iz);
            vx_make_tap(ix, iy, iz, g.X, g.Y, g.Z, t);
            const float fm = vx_tap_eval(grid, t);
            const float cp = sdf_tap_coords(g, s, a, disp.d[l], ix, iy, iz);
            vx_make_tap(ix, iy, iz, g.X, g.Y, g.Z, t);
            const float fp = vx_tap_eval(grid, t);
            gr[a] = __fdiv_rn(__fdiv_rn(__fsub_rn(fp, fm), __fsub_rn(cp, cm)), voxel_size);
          }
          const float nrm = sqrtf(gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2]);
          const float den = nrm + 1e-5f;
          const float dot = dgr[0] * gr[0] + dgr[1] * gr[1] + dgr[2] * gr[2];
          const float k = (nrm > 0.f) ? dot / (den * den * nrm) : 0.f;
#pragma unroll
          for (int a = 0; a < 3; ++a) dgr[a] = dgr[a] / den - gr[a] * k;
        }
      }
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const int aa = xyz_order ? (2 - a) : a;
        float dfm = grad_feat ? grad_feat[p * 6 * L + (aa * 2 + 0) * L + l] : 0.f;
        float dfp = grad_feat ? grad_feat[p * 6 * L + (aa * 2 + 1) * L + l] : 0.f;
        float ix, iy, iz;
        const float cm = sdf_tap_coords(g, s, a, -disp.d[l], ix, iy, iz);
        VxTap tm;
        vx_make_tap(ix, iy, iz, g.X, g.Y, g.Z, tm);
        const float cp = sdf_tap_coords(g, s, a, disp.d[l], ix, iy, iz);
        vx_make_tap(ix, iy, iz, g.X, g.Y, g.Z, t);
        if (dgr[a] != 0.f) {
          const float d = (dgr[a] / voxel_size) / (cp - cm);
          dfp += d;
          dfm -= d;
        }
        vx_tap_scatter(grad_grid, tm, dfm);
        vx_tap_scatter(grad_grid, t, dfp);
      }
    }
  }
}

VX_API int vx_sdf_taps(const float* grid, int X, int Y, int Z, const float* xyz_min_host, const float* xyz_max_host,
                       const float* xyz, const int* ray_id, const int* step_id, const float* rays_start,
                       const float* rays_dir, float stepdist, const int* n_dev, int64_t n_host,
                       const float* displace_host, int L, float voxel_size, int use_grad_norm, int xyz_order,
                       float* out_sdf, float* out_feat, float* out_grad, cudaStream_t st) {
  if (!n_dev && n_host <= 0) return 0;
  VxDisp d;
  VX_REQUIRE(fill_disp(d, displace_host, L) == 0, "vx_sdf_taps", "L out of range (0..8)");
  VX_REQUIRE(!xyz_order || L <= 1, "vx_sdf_taps", "xyz_order needs L == 1");
  const VxGrid g = make_grid(X, Y, Z, 1, 0, xyz_min_host, xyz_max_host);
  const VxPts pts{xyz, ray_id, step_id, rays_start, rays_dir, stepdist};
  k_sdf_taps<<<launch_blocks(n_dev, n_host), 256, 0, st>>>(g, grid, pts, n_dev, n_host, d, voxel_size, use_grad_norm,
                                                           xyz_order, out_sdf, out_feat, out_grad);
  return vx_check_launch("vx_sdf_taps");
}

VX_API int vx_sdf_taps_backward(const float* grid, int X, int Y, int Z, const float* xyz_min_host,
                                const float* xyz_max_host, const float* xyz, const int* ray_id, const int* step_id,
                                const float* rays_start, const float* rays_dir, float stepdist, const int* n_dev,
                                int64_t n_host, const float* displace_host, int L, float voxel_size, int use_grad_norm,
                                int xyz_order, const float* grad_sdf, const float* grad_feat, const float* grad_grad,
                                float* grad_grid, cudaStream_t st) {
  if (!n_dev && n_host <= 0) return 0;
  VxDisp d;
  VX_REQUIRE(fill_disp(d, displace_host, L) == 0, "vx_sdf_taps_backward", "L out of range (0..8)");
  const VxGrid g = make_grid(X, Y, Z, 1, 0, xyz_min_host, xyz_max_host);
  const VxPts pts{xyz, ray_id, step_id, rays_start, rays_dir, stepdist};
  k_sdf_taps_bwd<<<launch_blocks(n_dev, n_host), 256, 0, st>>>(g, grid, pts, n_dev, n_host, d, voxel_size, use_grad_norm,
                                                               xyz_order, grad_sdf, grad_feat, grad_grad, grad_grid);
  return vx_check_launch("vx_sdf_taps_backward");
}

// ---------------------------------------------------------------------------------------------
// Mesh field query (lib/voxurf_fine.py:894-910 + lib/dvgo_ori.py:679-693): the trilinear value of a single-channel grid,
// optionally negated (extract_geometry queries -sdf), and optionally its 6-tap gradient (Voxurf.grid_sampler with
// sample_grad=True, displace 1.0: lib/voxurf_fine.py:502-534) on the lattice xs[i] x ys[j] x zs[k] -- the three
// torch.linspace axes of extract_fields, xs already restricted to an X-slab.  One thread per lattice point; the
// arithmetic per point is k_sdf_taps's (L = 1, xyz order), so results are bit-identical to the chunked
// meshgrid -> grid_sampler path, without materialising the (P,3) points.
// ---------------------------------------------------------------------------------------------
__global__ void k_sdf_lattice(VxGrid g, const float* __restrict__ grid, const float* __restrict__ xs, const float* __restrict__ ys,
                              const float* __restrict__ zs, int nx, int ny, int nz, float voxel_size, float sign,
                              float* __restrict__ out_u, float* __restrict__ out_grad) {
  const int64_t n = (int64_t)nx * ny * nz;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(p % nz), j = (int)((p / nz) % ny), i = (int)(p / ((int64_t)nz * ny));
    const float px = __ldg(xs + i), py = __ldg(ys + j), pz = __ldg(zs + k);
    VxTap t;
    {
      float ix, iy, iz;
      point_to_index(g, px, py, pz, ix, iy, iz);
      vx_make_tap(ix, iy, iz, g.X, g.Y, g.Z, t);
      out_u[p] = sign * vx_tap_eval(grid, t);
    }
    if (!out_grad) continue;
    SdfTapCoords s;
    sdf_tap_setup(g, px, py, pz, s);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float ix, iy, iz;
      const float cm = sdf_tap_coords(g, s, a, -1.f, ix, iy, iz);
      vx_make_tap(ix, iy, iz, g.X, g.Y, g.Z, t);
      const float fm = vx_tap_eval(grid, t);
      const float cp = sdf_tap_coords(g, s, a, 1.f, ix, iy, iz);
      vx_make_tap(ix, iy, iz, g.X, g.Y, g.Z, t);
      const float fp = vx_tap_eval(grid, t);
      out_grad[p * 3 + (2 - a)] = __fdiv_rn(__fdiv_rn(__fsub_rn(fp, fm), __fsub_rn(cp, cm)), voxel_size);
    }
  }
}

VX_API int vx_sdf_lattice(const float* grid, int X, int Y, int Z, const float* xyz_min_host, const float* xyz_max_host,
                          const float* xs, const float* ys, const float* zs, int nx, int ny, int nz, float voxel_size,
                          int negate, float* out_u, float* out_grad, cudaStream_t st) {
  const int64_t n = (int64_t)nx * ny * nz;
  if (n <= 0) return 0;
  VX_REQUIRE(grid && xs && ys && zs && out_u, "vx_sdf_lattice", "null pointer");
  const VxGrid g = make_grid(X, Y, Z, 1, 0, xyz_min_host, xyz_max_host);
  k_sdf_lattice<<<(int)min((int64_t)vx_blocks(n, 256), (int64_t)vx_num_sms() * 32), 256, 0, st>>>(
      g, grid, xs, ys, zs, nx, ny, nz, voxel_size, negate ? -1.f : 1.f, out_u, out_grad);
  return vx_check_launch("vx_sdf_lattice");
}

// ---------------------------------------------------------------------------------------------
// NeuS alpha (lib/voxurf_fine.py:463-500), forward and backward, one thread per sample.
// ---------------------------------------------------------------------------------------------
__global__ void k_neus_alpha(const float* __restrict__ viewdirs, const int* __restrict__ ray_id,
                             const int64_t* __restrict__ ray_id64, const float* __restrict__ sdf,
                             const float* __restrict__ gradient, float dist, float inv_s, const int* __restrict__ n_dev,
                             int64_t n_host, float* __restrict__ alpha, const float* __restrict__ inv_s_dev) {
  const int64_t n = vx_count(n_dev, n_host);
  if (inv_s_dev) inv_s = __ldg(inv_s_dev);     // CUDA-graph replays: the step-dependent 1/s lives in device memory
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = ray_id ? (int64_t)ray_id[i] : ray_id64[i];
    const float gx = gradient[3 * i], gy = gradient[3 * i + 1], gz = gradient[3 * i + 2];
    const float true_cos = __fadd_rn(__fadd_rn(__fmul_rn(viewdirs[3 * r], gx), __fmul_rn(viewdirs[3 * r + 1], gy)),
                                     __fmul_rn(viewdirs[3 * r + 2], gz));
    const float iter_cos = -fmaxf(-true_cos, 0.f);
    const float h = __fmul_rn(__fmul_rn(iter_cos, dist), 0.5f);
    const float prev = sigmoidf_(__fmul_rn(__fsub_rn(sdf[i], h), inv_s));
    const float next = sigmoidf_(__fmul_rn(__fadd_rn(sdf[i], h), inv_s));
    const float a = __fdiv_rn(__fadd_rn(__fsub_rn(prev, next), 1e-5f), __fadd_rn(prev, 1e-5f));
    alpha[i] = fminf(fmaxf(a, 0.f), 1.f);
  }
}

__global__ void k_neus_alpha_bwd(const float* __restrict__ viewdirs, const int* __restrict__ ray_id,
                                 const int64_t* __restrict__ ray_id64, const float* __restrict__ sdf,
                                 const float* __restrict__ gradient, float dist, float inv_s,
                                 const int* __restrict__ n_dev, int64_t n_host, const float* __restrict__ grad_alpha,
                                 int accumulate, float* __restrict__ grad_sdf, float* __restrict__ grad_gradient,
                                 const float* __restrict__ inv_s_dev) {
  const int64_t n = vx_count(n_dev, n_host);
  if (inv_s_dev) inv_s = __ldg(inv_s_dev);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float ga = grad_alpha[i];
    float ds = 0.f, dgx = 0.f, dgy = 0.f, dgz = 0.f;
    if (ga != 0.f) {
      const int64_t r = ray_id ? (int64_t)ray_id[i] : ray_id64[i];
      const float vx = viewdirs[3 * r], vy = viewdirs[3 * r + 1], vz = viewdirs[3 * r + 2];
      const float gx = gradient[3 * i], gy = gradient[3 * i + 1], gz = gradient[3 * i + 2];
      const float true_cos = __fadd_rn(__fadd_rn(__fmul_rn(vx, gx), __fmul_rn(vy, gy)), __fmul_rn(vz, gz));
      const float iter_cos = -fmaxf(-true_cos, 0.f);
      const float h = __fmul_rn(__fmul_rn(iter_cos, dist), 0.5f);
      const float prev = sigmoidf_(__fmul_rn(__fsub_rn(sdf[i], h), inv_s));
      const float next = sigmoidf_(__fmul_rn(__fadd_rn(sdf[i], h), inv_s));
      const float u = __fadd_rn(__fsub_rn(prev, next), 1e-5f);
      const float c = __fadd_rn(prev, 1e-5f);
      const float a = __fdiv_rn(u, c);
      if (a >= 0.f && a <= 1.f) {  // clip backward passes inside [0,1] (boundaries included, like torch.clamp)
        const float du = ga / c;                 // d alpha / d u
        const float dc = -ga * u / (c * c);      // d alpha / d c
        const float dprev = du + dc;
        const float dnext = -du;
        const float dep = dprev * prev * (1.f - prev) * inv_s;  // wrt est_prev
        const float den = dnext * next * (1.f - next) * inv_s;  // wrt est_next
        ds = dep + den;
        const float dh = den - dep;                              // est_next = sdf + h, est_prev = sdf - h
        const float dcos = (true_cos < 0.f) ? dh * dist * 0.5f : 0.f;
        dgx = dcos * vx; dgy = dcos * vy; dgz = dcos * vz;
      }
    }
    if (accumulate) {
      grad_sdf[i] += ds;
      grad_gradient[3 * i] += dgx; grad_gradient[3 * i + 1] += dgy; grad_gradient[3 * i + 2] += dgz;
    } else {
      grad_sdf[i] = ds;
      grad_gradient[3 * i] = dgx; grad_gradient[3 * i + 1] = dgy; grad_gradient[3 * i + 2] = dgz;
    }
  }
}

VX_API int vx_neus_alpha(const float* viewdirs, const int* ray_id, const int64_t* ray_id64, const float* sdf,
                         const float* gradient, float dist, float inv_s, const int* n_dev, int64_t n_host, float* alpha, const float* inv_s_dev,
                         cudaStream_t st) {
  if (!n_dev && n_host <= 0) return 0;
  VX_REQUIRE((ray_id != nullptr) != (ray_id64 != nullptr), "vx_neus_alpha", "give exactly one of ray_id / ray_id64");
  k_neus_alpha<<<launch_blocks(n_dev, n_host), 256, 0, st>>>(viewdirs, ray_id, ray_id64, sdf, gradient, dist, inv_s, n_dev,
                                                             n_host, alpha, inv_s_dev);
  return vx_check_launch("vx_neus_alpha");
}

VX_API int vx_neus_alpha_backward(const float* viewdirs, const int* ray_id, const int64_t* ray_id64, const float* sdf,
                                  const float* gradient, float dist, float inv_s, const int* n_dev, int64_t n_host,
                                  const float* grad_alpha, int accumulate, float* grad_sdf, float* grad_gradient,
                                  const float* inv_s_dev, cudaStream_t st) {
  if (!n_dev && n_host <= 0) return 0;
  VX_REQUIRE((ray_id != nullptr) != (ray_id64 != nullptr), "vx_neus_alpha_backward", "give exactly one of ray_id / ray_id64");
  k_neus_alpha_bwd<<<launch_blocks(n_dev, n_host), 256, 0, st>>>(viewdirs, ray_id, ray_id64, sdf, gradient, dist, inv_s,
                                                                 n_dev, n_host, grad_alpha, accumulate, grad_sdf,
                                                                 grad_gradient, inv_s_dev);
  return vx_check_launch("vx_neus_alpha_backward");
}

// ---------------------------------------------------------------------------------------------
// segment_coo(src, index, out, reduce='sum') with a sorted index (lib/voxurf_fine.py:753-777):
// the thread that sits on a segment head walks its segment in index order and adds the K sums into
// out[index] -- deterministic, no atomics, same accumulation order as a sequential index_add_.
// ---------------------------------------------------------------------------------------------
__global__ void k_segment_coo(const float* __restrict__ src, const int64_t* __restrict__ index, int64_t M, int K,
                              float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int64_t r = index[i];
  if (i > 0 && index[i - 1] == r) return;
  for (int k0 = 0; k0 < K; k0 += 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int kn = min(4, K - k0);
    for (int k = 0; k < kn; ++k) acc[k] = out[r * K + k0 + k];
    for (int64_t j = i; j < M && index[j] == r; ++j)
      for (int k = 0; k < kn; ++k) acc[k] += src[j * K + k0 + k];
    for (int k = 0; k < kn; ++k) out[r * K + k0 + k] = acc[k];
  }
}

VX_API int vx_segment_coo_sum(const float* src, const int64_t* index, int64_t M, int K, float* out, cudaStream_t st) {
  if (M <= 0 || K <= 0) return 0;
  k_segment_coo<<<vx_blocks(M, 256), 256, 0, st>>>(src, index, M, K, out);
  return vx_check_launch("vx_segment_coo_sum");
}
