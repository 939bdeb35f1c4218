// voxurf_b200 -- shared device helpers.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define VX_API extern "C" __attribute__((visibility("default")))

// error plumbing: every entry point returns 0 or a non-zero code; vx_last_error() gives the text.
void vx_set_error(const char* where, const char* what);
int vx_check_launch(const char* where);
#define VX_REQUIRE(cond, where, what) \
  do {                                \
    if (!(cond)) {                    \
      vx_set_error(where, what);      \
      return -1;                      \
    }                                 \
  } while (0)

static inline int vx_blocks(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }
int vx_num_sms();

// ---------------------------------------------------------------------------------------------
// Grid geometry.  A dense voxel grid of world size (X,Y,Z) spanning [xyz_min, xyz_max].
// `cl` selects the channel layout in memory: 0 = channel-major (C,X,Y,Z) like the reference's
// (1,C,X,Y,Z) tensor, 1 = channels-last (X,Y,Z,C) (torch.channels_last_3d), which turns the 8
// corner reads of a C-channel gather into 8 contiguous C-float vectors.
// ---------------------------------------------------------------------------------------------
struct VxGrid {
  int X, Y, Z, C, cl;
  float min[3];
  float max[3];
};

// world coordinate -> grid_sample-normalised coordinate, spelled exactly like the reference's
// ((xyz - xyz_min) / (xyz_max - xyz_min)) * 2 - 1   (lib/grid.py:52): one rounding per torch op.
__device__ __forceinline__ float vx_norm_coord(float p, float lo, float hi) {
  const float n = __fdiv_rn(__fsub_rn(p, lo), __fsub_rn(hi, lo));
  return __fsub_rn(__fmul_rn(n, 2.f), 1.f);
}
// ATen grid_sampler_unnormalize, align_corners=True: ((coord + 1) / 2) * (size - 1)
__device__ __forceinline__ float vx_unnorm_coord(float n, int size) {
  return __fmul_rn(__fmul_rn(__fadd_rn(n, 1.f), 0.5f), (float)(size - 1));
}

// One trilinear tap: the 8 corner offsets (-1 when outside: zeros padding) and weights, in ATen's
// corner order (tnw, tne, tsw, tse, bnw, bne, bsw, bse) with its weight formulas
// (aten/native/cuda/GridSampler.cu, grid_sampler_3d_kernel).  ix indexes Z (fastest), iy Y, iz X.
struct VxTap {
  int off[8];   // voxel linear index (xi*Y + yi)*Z + zi, or -1
  float w[8];
};

__device__ __forceinline__ void vx_make_tap(float ix, float iy, float iz, int X, int Y, int Z, VxTap& t) {
  const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;   // x0 along Z, y0 along Y, z0 along X
  const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix;
  const float wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
  const float wz1 = iz - fz, wz0 = (fz + 1.f) - iz;
  const bool vx0 = (x0 >= 0) & (x0 < Z), vx1 = (x0 + 1 >= 0) & (x0 + 1 < Z);
  const bool vy0 = (y0 >= 0) & (y0 < Y), vy1 = (y0 + 1 >= 0) & (y0 + 1 < Y);
  const bool vz0 = (z0 >= 0) & (z0 < X), vz1 = (z0 + 1 >= 0) & (z0 + 1 < X);
  const int base = (z0 * Y + y0) * Z + x0;
  const int sY = Z, sX = Y * Z;
  t.w[0] = wx0 * wy0 * wz0; t.off[0] = (vx0 & vy0 & vz0) ? base : -1;
  t.w[1] = wx1 * wy0 * wz0; t.off[1] = (vx1 & vy0 & vz0) ? base + 1 : -1;
  t.w[2] = wx0 * wy1 * wz0; t.off[2] = (vx0 & vy1 & vz0) ? base + sY : -1;
  t.w[3] = wx1 * wy1 * wz0; t.off[3] = (vx1 & vy1 & vz0) ? base + sY + 1 : -1;
  t.w[4] = wx0 * wy0 * wz1; t.off[4] = (vx0 & vy0 & vz1) ? base + sX : -1;
  t.w[5] = wx1 * wy0 * wz1; t.off[5] = (vx1 & vy0 & vz1) ? base + sX + 1 : -1;
  t.w[6] = wx0 * wy1 * wz1; t.off[6] = (vx0 & vy1 & vz1) ? base + sX + sY : -1;
  t.w[7] = wx1 * wy1 * wz1; t.off[7] = (vx1 & vy1 & vz1) ? base + sX + sY + 1 : -1;
}

// single-channel tap evaluation, ATen accumulation order
__device__ __forceinline__ float vx_tap_eval(const float* __restrict__ g, const VxTap& t) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (t.off[c] >= 0) acc += __ldg(g + t.off[c]) * t.w[c];
  return acc;
}

// Scatter targets.  VxAccF: fp32 atomics straight into the gradient grid (the sum depends on the order the atomics land
// in: the last bits differ from run to run).  VxAccQ: 64-bit fixed-point accumulators -- integer addition is associative,
// so the sum is the same whatever the order: bit-reproducible gradients (SURVEY.md 8e "Determinism"; the reference
// inherits ATen's fp32 atomics and says so, run.py:279).  One accumulator unit is 1 / scale; with scale = 2^52 a
// contribution >= 4e-9 keeps its full fp32 mantissa and sums up to +-2048 fit.
struct VxAccF {
  float* p;
  __device__ __forceinline__ void add(int64_t i, float v) const { atomicAdd(p + i, v); }
};
struct VxAccQ {
  unsigned long long* p;
  double scale;
  __device__ __forceinline__ void add(int64_t i, float v) const {
    atomicAdd(p + i, (unsigned long long)__double2ll_rn((double)v * scale));
  }
};

// scatter g * w into grad at the tap's corners; exact zeros are skipped (adding +0 is a no-op)
template <class Acc>
__device__ __forceinline__ void vx_tap_scatter(const Acc& acc, const VxTap& t, float g) {
  if (g == 0.f) return;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (t.off[c] >= 0) acc.add(t.off[c], g * t.w[c]);
}
__device__ __forceinline__ void vx_tap_scatter(float* __restrict__ grad, const VxTap& t, float g) {
  vx_tap_scatter(VxAccF{grad}, t, g);
}

// Sample positions are either an explicit (P,3) array or implicit: (ray_id, step_id) into per-ray
// start/dir, p = start + dir * (stepdist * step) -- the reference's formula
// (lib/cuda/render_utils_kernel.cu:184-187) so implicit and explicit points are bit-identical.
struct VxPts {
  const float* xyz;        // (P,3) or nullptr
  const int* ray_id;       // (P,)
  const int* step_id;      // (P,)
  const float* rays_start; // (N,3)
  const float* rays_dir;   // (N,3)
  float stepdist;
};

__device__ __forceinline__ void vx_load_pt(const VxPts& s, int64_t i, float& px, float& py, float& pz) {
  if (s.xyz) {
    px = __ldg(s.xyz + 3 * i); py = __ldg(s.xyz + 3 * i + 1); pz = __ldg(s.xyz + 3 * i + 2);
  } else {
    const int r = __ldg(s.ray_id + i);
    const float dist = s.stepdist * __ldg(s.step_id + i);
    px = __ldg(s.rays_start + 3 * r) + __ldg(s.rays_dir + 3 * r) * dist;
    py = __ldg(s.rays_start + 3 * r + 1) + __ldg(s.rays_dir + 3 * r + 1) * dist;
    pz = __ldg(s.rays_start + 3 * r + 2) + __ldg(s.rays_dir + 3 * r + 2) * dist;
  }
}

// number of items for kernels whose size lives on the device (sync-free pipelines): n_dev may be
// nullptr, then n_host is used.
__device__ __forceinline__ int64_t vx_count(const int* n_dev, int64_t n_host) {
  return n_dev ? (int64_t)__ldg(n_dev) : n_host;
}
