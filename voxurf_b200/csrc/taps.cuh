// voxurf_b200 -- shared helpers for SDF taps (lib/voxurf_fine.py:537-577) and NeuS alpha (lib/voxurf_fine.py:463-500).
#pragma once
#include "common.cuh"

__device__ __forceinline__ void point_to_index(const VxGrid& g, float px, float py, float pz, float& ix, float& iy,
                                               float& iz) {
  iz = vx_unnorm_coord(vx_norm_coord(px, g.min[0], g.max[0]), g.X);  // world x -> slowest dim (ATen z / D)
  iy = vx_unnorm_coord(vx_norm_coord(py, g.min[1], g.max[1]), g.Y);
  ix = vx_unnorm_coord(vx_norm_coord(pz, g.min[2], g.max[2]), g.Z);  // world z -> fastest dim (ATen x / W)
}


#define VX_MAX_L 8

struct VxDisp {
  int L;
  float d[VX_MAX_L];
};

__device__ __forceinline__ float roundtrip(float a, int size) {
  const float sm1 = (float)(size - 1);
  const float n = __fsub_rn(__fmul_rn(__fdiv_rn(a, sm1), 2.f), 1.f);
  return vx_unnorm_coord(n, size);
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// axis a in reference order: 0 = z (fastest, ATen x), 1 = y, 2 = x (slowest, ATen z)
struct SdfTapCoords {
  float c[3];    // undisplaced clamped+roundtripped coordinate per axis (ATen x,y,z order)
  float raw[3];  // unclamped index per axis
};

__device__ __forceinline__ void sdf_tap_setup(const VxGrid& g, float px, float py, float pz, SdfTapCoords& s) {
  float ix, iy, iz;
  point_to_index(g, px, py, pz, ix, iy, iz);
  s.raw[0] = ix; s.raw[1] = iy; s.raw[2] = iz;
  s.c[0] = roundtrip(clampf(ix, 0.f, (float)(g.Z - 1)), g.Z);
  s.c[1] = roundtrip(clampf(iy, 0.f, (float)(g.Y - 1)), g.Y);
  s.c[2] = roundtrip(clampf(iz, 0.f, (float)(g.X - 1)), g.X);
}

__device__ __forceinline__ int axis_size(const VxGrid& g, int a) { return a == 0 ? g.Z : (a == 1 ? g.Y : g.X); }

// coordinates of tap (axis a, sign s in {-1,+1}, displacement d); returns the clamped raw index on axis a
__device__ __forceinline__ float sdf_tap_coords(const VxGrid& g, const SdfTapCoords& s, int a, float sd, float& ix,
                                                float& iy, float& iz) {
  const int size = axis_size(g, a);
  const float cl = clampf(__fadd_rn(s.raw[a], sd), 0.f, (float)(size - 1));
  const float v = roundtrip(cl, size);
  ix = (a == 0) ? v : s.c[0];
  iy = (a == 1) ? v : s.c[1];
  iz = (a == 2) ? v : s.c[2];
  return cl;
}


// one axis of a trilinear tap, spelled like vx_make_tap: floor index, the two weights, validity of index i0 / i0 + 1
struct AxisTap {
  int i0;
  float w0, w1;
  bool v0, v1;
};
__device__ __forceinline__ AxisTap axis_tap(float c, int size) {
  const float f = floorf(c);
  AxisTap a;
  a.i0 = (int)f;
  a.w1 = c - f;
  a.w0 = (f + 1.f) - c;
  a.v0 = (a.i0 >= 0) & (a.i0 < size);
  a.v1 = (a.i0 + 1 >= 0) & (a.i0 + 1 < size);
  return a;
}

// The three axis parts of a tap (az: Z / fastest, ay: Y, ax: X) -> value / scatter, with vx_make_tap's weights
// (wz * wy) * wx, corner order and zero padding, and vx_tap_eval's / vx_tap_scatter's accumulation.  The displaced taps
// of sample_sdfs differ from each other in one axis part only, so callers build the shared parts once.
__device__ __forceinline__ float tap_eval_axes(const float* __restrict__ grid, int Y, int Z, const AxisTap& az, const AxisTap& ay,
                                               const AxisTap& ax) {
  const int base = (ax.i0 * Y + ay.i0) * Z + az.i0, sY = Z, sX = Y * Z;
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int bz = c & 1, by = (c >> 1) & 1, bx = (c >> 2) & 1;
    const float w = (bz ? az.w1 : az.w0) * (by ? ay.w1 : ay.w0) * (bx ? ax.w1 : ax.w0);
    const bool valid = (bz ? az.v1 : az.v0) & (by ? ay.v1 : ay.v0) & (bx ? ax.v1 : ax.v0);
    if (valid) acc += __ldg(grid + base + bx * sX + by * sY + bz) * w;
  }
  return acc;
}

template <class Acc>
__device__ __forceinline__ void tap_scatter_axes(const Acc& acc, int Y, int Z, const AxisTap& az, const AxisTap& ay,
                                                 const AxisTap& ax, float g) {
  if (g == 0.f) return;   // adding +0 is a no-op
  const int base = (ax.i0 * Y + ay.i0) * Z + az.i0, sY = Z, sX = Y * Z;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int bz = c & 1, by = (c >> 1) & 1, bx = (c >> 2) & 1;
    const float w = (bz ? az.w1 : az.w0) * (by ? ay.w1 : ay.w0) * (bx ? ax.w1 : ax.w0);
    const bool valid = (bz ? az.v1 : az.v0) & (by ? ay.v1 : ay.v0) & (bx ? ax.v1 : ax.v0);
    if (valid) acc.add(base + bx * sX + by * sY + bz, g * w);
  }
}

// axis parts of the tap displaced by +-d along axis kA (0 = Z, 1 = Y, 2 = X): `disp` replaces the shared part of axis kA
#define VX_AXES(kA, disp, sz, sy, sx) ((kA) == 0 ? (disp) : (sz)), ((kA) == 1 ? (disp) : (sy)), ((kA) == 2 ? (disp) : (sx))


static inline VxGrid make_grid(int X, int Y, int Z, int C, int cl, const float* mn, const float* mx) {
  VxGrid g;
  g.X = X; g.Y = Y; g.Z = Z; g.C = C; g.cl = cl;
  for (int i = 0; i < 3; ++i) { g.min[i] = mn[i]; g.max[i] = mx[i]; }
  return g;
}

static inline int launch_blocks(const int* n_dev, int64_t n_host) {
  return n_dev ? vx_num_sms() * 8 : (int)min((int64_t)vx_blocks(n_host, 256), (int64_t)vx_num_sms() * 16);
}


static inline int fill_disp(VxDisp& d, const float* displace_host, int L) {
  if (L < 0 || L > VX_MAX_L) return -1;
  d.L = L;
  for (int i = 0; i < VX_MAX_L; ++i) d.d[i] = i < L ? displace_host[i] : 0.f;
  return 0;
}


__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
