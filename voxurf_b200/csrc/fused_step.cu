// Fused, sync-free fine-stage step: the kernels between the
// ray march (sampling.cu) and the optimizer (adam.cu) of one Voxurf.forward + loss + backward
// (lib/voxurf_fine.py:620-802, run.py:604-639).
//
// Layout in HBM (all persistent, capacity sized once; nothing is allocated per step):
//   M2-level arrays ("samples that survived bbox + mask cache", ray-sorted, ~2 M):
//     ray_id, step_id (int32), sdf, grad[3], alpha, keep (u8: alpha > thres), weight, T,
//     d_w, d_alpha, d_sdf_s, d_grad_s[3]
//   M4-level arrays ("rows": samples with weight > thres, ray-sorted, ~0.06-0.16 M):
//     idx4 (row -> M2 index), X1 [rows, ld1] rgbnet input, X2 [rows, ld2] k_rgbnet input
//   per-ray arrays: keep_off (M2 segment offsets), off4 (row segment offsets), alphainv_last, i_end
// Every count lives on the device (keep_off[N] = M2, off4[N] = M4): no host round trip, the whole step is
// CUDA-graph capturable.  Threshold compactions of the reference (6-7 boolean-index gathers each,
// lib/voxurf_fine.py:647-676) become a keep flag (first threshold) and one index list (second threshold).
#include "common.cuh"
#include "taps.cuh"

// ---------------------------------------------------------------------------------------------
// S4: SDF value + 6-neighbour gradient + NeuS alpha for every M2 sample
//     (grid_sampler(sample_grad=True) lib/voxurf_fine.py:640 + neus_alpha_from_sdf_scatter :643 + mask :648)
// ---------------------------------------------------------------------------------------------
// The 7 trilinear taps of a sample (centre + one voxel to either side on each axis) read 56 corners, but only 32
// distinct voxels: the centre cell plus one more layer on each of its six faces.  The kernel is bound by L1 wavefronts
// (every corner read of a warp spreads over ~10 sectors), so the 32 voxels are loaded once into registers and every tap
// whose cell is where it is expected -- all of them, except when a coordinate sits within an ulp of a cell boundary or
// is clamped at the grid border -- takes its corners from there; the arithmetic (weights, corner order, predicates)
// is vx_tap_eval's, so the result is bit-identical to evaluating each tap from memory, which remains the fallback.
struct SdfNeighbourhood {
  float cube[8];        // corner c of the centre cell (bit 0: +Z, bit 1: +Y, bit 2: +X)
  float ext[3][2][4];   // [axis][0: index -1, 1: index +2][the other two axes' bits, lower axis first]
};

__device__ __forceinline__ void nb_fill(const float* __restrict__ grid, const VxGrid& g, int cx, int cy, int cz,
                                        SdfNeighbourhood& nb) {
  // validity of the cell indices -1..2 on each axis, once; every voxel then costs two ANDs and an add
  bool okx[4], oky[4], okz[4];
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    okx[o] = (cx + o - 1 >= 0) & (cx + o - 1 < g.X);
    oky[o] = (cy + o - 1 >= 0) & (cy + o - 1 < g.Y);
    okz[o] = (cz + o - 1 >= 0) & (cz + o - 1 < g.Z);
  }
  const int sY = g.Z, sX = g.Y * g.Z;
  const float* base = grid + ((int64_t)cx * g.Y + cy) * g.Z + cz;
  auto ld = [&](int dx, int dy, int dz) -> float {   // offsets -1..2, compile-time after unrolling
    return (okx[dx + 1] & oky[dy + 1] & okz[dz + 1]) ? __ldg(base + dx * sX + dy * sY + dz) : 0.f;
  };
#pragma unroll
  for (int c = 0; c < 8; ++c) nb.cube[c] = ld((c >> 2) & 1, (c >> 1) & 1, c & 1);
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    const int o = m ? 2 : -1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      nb.ext[0][m][i] = ld((i >> 1) & 1, i & 1, o);        // Z displaced: bits (Y, X)
      nb.ext[1][m][i] = ld((i >> 1) & 1, o, i & 1);        // Y displaced: bits (Z, X)
      nb.ext[2][m][i] = ld(o, (i >> 1) & 1, i & 1);        // X displaced: bits (Z, Y)
    }
  }
}

// Trilinear tap from its three axis parts (az: Z / fastest, ay: Y, ax: X), displaced by kS cells along axis kA relative
// to the neighbourhood's anchor cell (kS = 0: the anchor cell itself).  Weights (wz * wy) * wx, corner order, the
// validity predicates and the accumulation are those of vx_make_tap + vx_tap_eval.
template <int kA, int kS>
__device__ __forceinline__ float nb_tap_eval(const SdfNeighbourhood& nb, const AxisTap& az, const AxisTap& ay, const AxisTap& ax) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int bz = c & 1, by = (c >> 1) & 1, bx = (c >> 2) & 1;
    const float w = (bz ? az.w1 : az.w0) * (by ? ay.w1 : ay.w0) * (bx ? ax.w1 : ax.w0);
    const bool valid = (bz ? az.v1 : az.v0) & (by ? ay.v1 : ay.v0) & (bx ? ax.v1 : ax.v0);
    const int bit = (c >> kA) & 1;
    const int pos = kS + bit;                                   // cell index on the displaced axis: -1..2
    const int oi = (kA == 0) ? (c >> 1) : (kA == 1) ? ((c & 1) | ((c >> 1) & 2)) : (c & 3);
    const float v = (pos == -1) ? nb.ext[kA][0][oi] : (pos == 2) ? nb.ext[kA][1][oi] : nb.cube[(c & ~(1 << kA)) | (pos << kA)];
    if (valid) acc += v * w;
  }
  return acc;
}

// value of the tap at index coordinates (ix, iy, iz) whose axis parts are (az, ay, ax); from the neighbourhood when the
// tap's cell is the expected one, from memory otherwise
template <int kA, int kS>
__device__ __forceinline__ float sdf_tap_value(const VxGrid& g, const float* __restrict__ grid, const SdfNeighbourhood& nb,
                                               int cx, int cy, int cz, const AxisTap& az, const AxisTap& ay, const AxisTap& ax,
                                               float ix, float iy, float iz) {
  const bool cached = (ax.i0 == cx + (kA == 2 ? kS : 0)) & (ay.i0 == cy + (kA == 1 ? kS : 0)) & (az.i0 == cz + (kA == 0 ? kS : 0));
  if (cached) return nb_tap_eval<kA, kS>(nb, az, ay, ax);
  VxTap t;
  vx_make_tap(ix, iy, iz, g.X, g.Y, g.Z, t);
  return vx_tap_eval(grid, t);
}

// (f(+1) - f(-1)) / (c(+1) - c(-1)) / voxel_size along axis kA; sz / sy / sx: the axis parts of the shared, undisplaced
// (clamped, round-tripped) coordinates tc.c[]
template <int kA>
__device__ __forceinline__ float sdf_axis_gradient(const VxGrid& g, const float* __restrict__ grid, const SdfNeighbourhood& nb,
                                                   const SdfTapCoords& tc, int cx, int cy, int cz, const AxisTap& sz,
                                                   const AxisTap& sy, const AxisTap& sx, float voxel_size) {
  float ix, iy, iz;
  const int size = axis_size(g, kA);
  const float cm = sdf_tap_coords(g, tc, kA, -1.f, ix, iy, iz);
  const AxisTap dm = axis_tap(kA == 0 ? ix : (kA == 1 ? iy : iz), size);
  const float fm = sdf_tap_value<kA, -1>(g, grid, nb, cx, cy, cz, kA == 0 ? dm : sz, kA == 1 ? dm : sy, kA == 2 ? dm : sx, ix, iy, iz);
  const float cp = sdf_tap_coords(g, tc, kA, 1.f, ix, iy, iz);
  const AxisTap dp = axis_tap(kA == 0 ? ix : (kA == 1 ? iy : iz), size);
  const float fp = sdf_tap_value<kA, 1>(g, grid, nb, cx, cy, cz, kA == 0 ? dp : sz, kA == 1 ? dp : sy, kA == 2 ? dp : sx, ix, iy, iz);
  return __fdiv_rn(__fdiv_rn(__fsub_rn(fp, fm), __fsub_rn(cp, cm)), voxel_size);
}

__global__ void __launch_bounds__(256) k_sdf_alpha_fwd(VxGrid g, const float* __restrict__ grid, VxPts pts, const int* __restrict__ n_dev,
                                const float* __restrict__ viewdirs, float voxel_size, float dist, float inv_s,
                                float thres, float* __restrict__ sdf, float* __restrict__ grad,
                                float* __restrict__ alpha, uint8_t* __restrict__ keep, float* __restrict__ d_w,
                                float* __restrict__ d_sdf_s, float* __restrict__ d_grad_s,
                                const float* __restrict__ inv_s_dev) {
  const int64_t n = *n_dev;
  if (inv_s_dev) inv_s = __ldg(inv_s_dev);   // CUDA-graph replays: the step-dependent 1/s lives in device memory
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    float px, py, pz;
    vx_load_pt(pts, p, px, py, pz);
    float ix, iy, iz;
    point_to_index(g, px, py, pz, ix, iy, iz);
    SdfTapCoords tc;
    sdf_tap_setup(g, px, py, pz, tc);
    // neighbourhood anchored at the cell of the displaced taps' shared (clamped, round-tripped) coordinates; the axis
    // parts of those coordinates are shared by the six displaced taps (each replaces one of them)
    const AxisTap sz = axis_tap(tc.c[0], g.Z), sy = axis_tap(tc.c[1], g.Y), sx = axis_tap(tc.c[2], g.X);
    const int cz = sz.i0, cy = sy.i0, cx = sx.i0;
    SdfNeighbourhood nb;
    nb_fill(grid, g, cx, cy, cz, nb);
    const float s = sdf_tap_value<0, 0>(g, grid, nb, cx, cy, cz, axis_tap(ix, g.Z), axis_tap(iy, g.Y), axis_tap(iz, g.X), ix, iy, iz);
    // reference axis order z,y,x
    const float gz = sdf_axis_gradient<0>(g, grid, nb, tc, cx, cy, cz, sz, sy, sx, voxel_size);
    const float gy = sdf_axis_gradient<1>(g, grid, nb, tc, cx, cy, cz, sz, sy, sx, voxel_size);
    const float gx = sdf_axis_gradient<2>(g, grid, nb, tc, cx, cy, cz, sz, sy, sx, voxel_size);
    const int r = pts.ray_id[p];
    const float true_cos = __fadd_rn(__fadd_rn(__fmul_rn(viewdirs[3 * r], gx), __fmul_rn(viewdirs[3 * r + 1], gy)),
                                     __fmul_rn(viewdirs[3 * r + 2], gz));
    const float iter_cos = -fmaxf(-true_cos, 0.f);
    const float h = __fmul_rn(__fmul_rn(iter_cos, dist), 0.5f);
    const float prev = sigmoidf_(__fmul_rn(__fsub_rn(s, h), inv_s));
    const float next = sigmoidf_(__fmul_rn(__fadd_rn(s, h), inv_s));
    float a_ = __fdiv_rn(__fadd_rn(__fsub_rn(prev, next), 1e-5f), __fadd_rn(prev, 1e-5f));
    a_ = fminf(fmaxf(a_, 0.f), 1.f);
    sdf[p] = s;
    grad[3 * p] = gx; grad[3 * p + 1] = gy; grad[3 * p + 2] = gz;
    alpha[p] = a_;
    keep[p] = a_ > thres;
    d_w[p] = 0.f; d_sdf_s[p] = 0.f;
    d_grad_s[3 * p] = 0.f; d_grad_s[3 * p + 1] = 0.f; d_grad_s[3 * p + 2] = 0.f;
  }
}

VX_API int vx_fused_sdf_alpha(const float* grid, int X, int Y, int Z, const float* xyz_min_host, const float* xyz_max_host,
                              const int* ray_id, const int* step_id, const float* rays_start, const float* rays_dir,
                              float stepdist, const int* n_dev, const float* viewdirs, float voxel_size, float dist,
                              float inv_s, float thres, float* sdf, float* grad, float* alpha, uint8_t* keep, float* d_w,
                              float* d_sdf_s, float* d_grad_s, const float* inv_s_dev, cudaStream_t st) {
  VX_REQUIRE(n_dev != nullptr, "vx_fused_sdf_alpha", "n_dev required");
  const VxGrid g = make_grid(X, Y, Z, 1, 0, xyz_min_host, xyz_max_host);
  const VxPts pts{nullptr, ray_id, step_id, rays_start, rays_dir, stepdist};
  k_sdf_alpha_fwd<<<vx_num_sms() * 8, 256, 0, st>>>(g, grid, pts, n_dev, viewdirs, voxel_size, dist, inv_s, thres, sdf,
                                                    grad, alpha, keep, d_w, d_sdf_s, d_grad_s, inv_s_dev);
  return vx_check_launch("vx_fused_sdf_alpha");
}

// ---------------------------------------------------------------------------------------------
// S6: rows = samples with weight > thres, in ray order (lib/voxurf_fine.py:668-676)
// ---------------------------------------------------------------------------------------------
__global__ void k_emit_rows(const uint8_t* __restrict__ w_keep, const int* __restrict__ seg_off, const int* __restrict__ off4,
                            int n_rays, int capacity, int* __restrict__ idx4, int* __restrict__ overflow) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * warps_per_block) {
    const int i_s = seg_off[r], i_e = seg_off[r + 1];
    int out = off4[r];
    for (int base = i_s; base < i_e; base += 32) {
      const int i = base + lane;
      const bool f = (i < i_e) && w_keep[i];
      const uint32_t m = __ballot_sync(0xffffffffu, f);
      if (f) {
        const int dst = out + __popc(m & ((1u << lane) - 1u));
        if (dst < capacity) idx4[dst] = i;
      }
      out += __popc(m);
    }
    if (r == n_rays - 1 && lane == 0 && out > capacity) *overflow = out;
  }
}

VX_API int vx_fused_emit_rows(const uint8_t* w_keep, const int* seg_off, const int* off4, int n_rays, int capacity,
                              int* idx4, int* overflow, cudaStream_t st) {
  if (n_rays <= 0) return 0;
  const int blocks = min(vx_blocks((int64_t)n_rays * 32, 256), vx_num_sms() * 8);
  k_emit_rows<<<blocks, 256, 0, st>>>(w_keep, seg_off, off4, n_rays, capacity, idx4, overflow);
  return vx_check_launch("vx_fused_emit_rows");
}

// ---------------------------------------------------------------------------------------------
// S7: MLP input rows (lib/voxurf_fine.py:678-739).  Column layout (P = posbase_pe, Vp = viewbase_pe, L):
//   X1: [xyz 3 | sin 3P | cos 3P | view 3 | sin 3Vp | cos 3Vp | sdf 1 | all_feat 6L | all_grad 3L]      = D1
//   X2: [k0 C | xyz 3 | sin 3P2 | cos 3P2 | view 3 | sin 3V2 | cos 3V2 | gradient 3 | rgb_logit 3 (later)] = D2
// Rows >= M4 (up to `capacity`) are zero-filled so a fixed-shape GEMM over the capacity is harmless.
// ---------------------------------------------------------------------------------------------
struct VxRowLayout {
  int P, Vp, P2, V2, L, C;
  int ld1, ld2;   // row strides (>= D1, D2)
  float disp[VX_MAX_L];
};

template <int kC>
__global__ void k_row_features(VxGrid gs, const float* __restrict__ sdf_grid, VxGrid gk, const float* __restrict__ k0_grid,
                               VxPts pts, const int* __restrict__ idx4, const int* __restrict__ n_rows_dev, int capacity,
                               const float* __restrict__ viewdirs, const float* __restrict__ sdf_s,
                               const float* __restrict__ grad_s, float voxel_size, int use_grad_norm, VxRowLayout lay,
                               float* __restrict__ X1, float* __restrict__ X2) {
  // four threads per row: sub-thread s evaluates the displacements l = s, s+4, ... (12 taps each); sub-thread 0 also
  // writes the positional encodings, sub-thread 1 the k0 gather, sub-thread 2 the centre sdf / gradient columns
  const int n = min(*n_rows_dev, capacity);
  for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < capacity * 4; item += gridDim.x * blockDim.x) {
    const int row = item >> 2, sub = item & 3;
    float* x1 = X1 + (int64_t)row * lay.ld1;
    float* x2 = X2 + (int64_t)row * lay.ld2;
    if (row >= n) {
      for (int c = sub; c < lay.ld1; c += 4) x1[c] = 0.f;
      for (int c = sub; c < lay.ld2; c += 4) x2[c] = 0.f;
      continue;
    }
    const int i = idx4[row];
    float p[3];
    vx_load_pt(pts, i, p[0], p[1], p[2]);
    const int r = pts.ray_id[i];
    // ---- positional encodings (lib/voxurf_fine.py:694-698, 723-726)
    float xn[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) xn[d] = __fdiv_rn(__fsub_rn(p[d], gs.min[d]), __fsub_rn(gs.max[d], gs.min[d]));
    int c1 = 0, c2 = kC;
    float vd[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) vd[d] = viewdirs[3 * r + d];
    // The sin / cos evaluations dominate this kernel: both networks take the same encodings of the same arguments, so
    // every (dimension, frequency) pair is evaluated once, written to both rows, and the pairs are dealt round-robin
    // to the four sub-threads of the row.
    if (sub == 0) {
#pragma unroll
      for (int d = 0; d < 3; ++d) { x1[c1 + d] = xn[d]; x2[c2 + d] = xn[d]; }
    }
    c1 += 3; c2 += 3;
    {
      const int Pm = max(lay.P, lay.P2);
      for (int q = sub; q < 3 * Pm; q += 4) {
        const int d = q / Pm, f = q - d * Pm;
        const float e = __fmul_rn(d == 0 ? xn[0] : (d == 1 ? xn[1] : xn[2]), (float)(1 << f));
        float sn, cs;
        sincosf(e, &sn, &cs);
        if (f < lay.P) { x1[c1 + d * lay.P + f] = sn; x1[c1 + 3 * lay.P + d * lay.P + f] = cs; }
        if (f < lay.P2) { x2[c2 + d * lay.P2 + f] = sn; x2[c2 + 3 * lay.P2 + d * lay.P2 + f] = cs; }
      }
    }
    c1 += 6 * lay.P; c2 += 6 * lay.P2;
    if (sub == 1) {
#pragma unroll
      for (int d = 0; d < 3; ++d) { x1[c1 + d] = vd[d]; x2[c2 + d] = vd[d]; }
    }
    c1 += 3; c2 += 3;
    {
      const int Vm = max(lay.Vp, lay.V2);
      for (int q = sub; q < 3 * Vm; q += 4) {
        const int d = q / Vm, f = q - d * Vm;
        const float e = __fmul_rn(d == 0 ? vd[0] : (d == 1 ? vd[1] : vd[2]), (float)(1 << f));
        float sn, cs;
        sincosf(e, &sn, &cs);
        if (f < lay.Vp) { x1[c1 + d * lay.Vp + f] = sn; x1[c1 + 3 * lay.Vp + d * lay.Vp + f] = cs; }
        if (f < lay.V2) { x2[c2 + d * lay.V2 + f] = sn; x2[c2 + 3 * lay.V2 + d * lay.V2 + f] = cs; }
      }
    }
    c1 += 6 * lay.Vp; c2 += 6 * lay.V2;
    // ---- centre sdf, gradient (values of the M2-level pass)
    if (sub == 2) x1[c1] = sdf_s[i];
    c1 += 1;
    if (sub == 2) { x2[c2] = grad_s[3 * i]; x2[c2 + 1] = grad_s[3 * i + 1]; x2[c2 + 2] = grad_s[3 * i + 2]; }
    c2 += 3;
    if (sub == 2) {
      x2[c2] = 0.f; x2[c2 + 1] = 0.f; x2[c2 + 2] = 0.f;   // rgb_logit.detach(), filled after the first MLP
      for (int c = c2 + 3; c < lay.ld2; ++c) x2[c] = 0.f;
    }
    // ---- sample_sdfs (displacement list, normalised gradients) lib/voxurf_fine.py:688
    const int L = lay.L;
    SdfTapCoords tc;
    sdf_tap_setup(gs, p[0], p[1], p[2], tc);
    const AxisTap sz = axis_tap(tc.c[0], gs.Z), sy = axis_tap(tc.c[1], gs.Y), sx = axis_tap(tc.c[2], gs.X);   // shared axis parts
    VxTap t;
    for (int l = sub; l < L; l += 4) {
      float gr[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        float ix, iy, iz;
        const int size = axis_size(gs, a);
        const float cm = sdf_tap_coords(gs, tc, a, -lay.disp[l], ix, iy, iz);
        const AxisTap dm = axis_tap(a == 0 ? ix : (a == 1 ? iy : iz), size);
        const float fm = tap_eval_axes(sdf_grid, gs.Y, gs.Z, VX_AXES(a, dm, sz, sy, sx));
        const float cp = sdf_tap_coords(gs, tc, a, lay.disp[l], ix, iy, iz);
        const AxisTap dp = axis_tap(a == 0 ? ix : (a == 1 ? iy : iz), size);
        const float fp = tap_eval_axes(sdf_grid, gs.Y, gs.Z, VX_AXES(a, dp, sz, sy, sx));
        gr[a] = __fdiv_rn(__fdiv_rn(__fsub_rn(fp, fm), __fsub_rn(cp, cm)), voxel_size);
        x1[c1 + (a * 2 + 0) * L + l] = fm;
        x1[c1 + (a * 2 + 1) * L + l] = fp;
      }
      if (use_grad_norm) {
        const float nrm = sqrtf(gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2]) + 1e-5f;
        gr[0] = gr[0] / nrm; gr[1] = gr[1] / nrm; gr[2] = gr[2] / nrm;
      }
#pragma unroll
      for (int a = 0; a < 3; ++a) x1[c1 + 6 * L + a * L + l] = gr[a];
    }
    if (sub == 3)
      for (int c = c1 + 9 * L; c < lay.ld1; ++c) x1[c] = 0.f;
    // ---- k0 trilinear gather (DenseGrid.forward lib/grid.py:47-58) into X2[:, 0:C]
    if (sub == 1) {
      float ix, iy, iz;
      point_to_index(gk, p[0], p[1], p[2], ix, iy, iz);
      vx_make_tap(ix, iy, iz, gk.X, gk.Y, gk.Z, t);
      const int64_t V = (int64_t)gk.X * gk.Y * gk.Z;
      float acc[kC];
#pragma unroll
      for (int c = 0; c < kC; ++c) acc[c] = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (t.off[k] < 0) continue;
        if (gk.cl) {
          const float* src = k0_grid + (int64_t)t.off[k] * kC;
          if (kC % 4 == 0) {
#pragma unroll
            for (int c = 0; c < kC; c += 4) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(src + c));
              acc[c] += v.x * t.w[k]; acc[c + 1] += v.y * t.w[k]; acc[c + 2] += v.z * t.w[k]; acc[c + 3] += v.w * t.w[k];
            }
          } else {
#pragma unroll
            for (int c = 0; c < kC; c += 2) {
              const float2 v = __ldg(reinterpret_cast<const float2*>(src + c));
              acc[c] += v.x * t.w[k]; acc[c + 1] += v.y * t.w[k];
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < kC; ++c) acc[c] += __ldg(k0_grid + c * V + t.off[k]) * t.w[k];
        }
      }
#pragma unroll
      for (int c = 0; c < kC; ++c) x2[c] = acc[c];
    }
  }
}

static int fill_layout(VxRowLayout& lay, int P, int Vp, int P2, int V2, int L, int C, int ld1, int ld2, const float* disp_host) {
  if (L < 0 || L > VX_MAX_L) return -1;
  lay.P = P; lay.Vp = Vp; lay.P2 = P2; lay.V2 = V2; lay.L = L; lay.C = C; lay.ld1 = ld1; lay.ld2 = ld2;
  for (int i = 0; i < VX_MAX_L; ++i) lay.disp[i] = i < L ? disp_host[i] : 0.f;
  const int D1 = 3 + 6 * P + 3 + 6 * Vp + 1 + 9 * L, D2 = C + 3 + 6 * P2 + 3 + 6 * V2 + 3 + 3;
  return (ld1 >= D1 && ld2 >= D2) ? 0 : -2;
}

VX_API int vx_fused_row_features(const float* sdf_grid, const float* k0_grid, int X, int Y, int Z, int C, int k0_channels_last,
                                 const float* xyz_min_host, const float* xyz_max_host, const int* ray_id, const int* step_id,
                                 const float* rays_start, const float* rays_dir, float stepdist, const int* idx4,
                                 const int* n_rows_dev, int capacity, const float* viewdirs, const float* sdf_s,
                                 const float* grad_s, float voxel_size, int use_grad_norm, int P, int Vp, int P2, int V2,
                                 const float* displace_host, int L, int ld1, int ld2, float* X1, float* X2,
                                 cudaStream_t st) {
  if (capacity <= 0) return 0;
  VxRowLayout lay;
  VX_REQUIRE(fill_layout(lay, P, Vp, P2, V2, L, C, ld1, ld2, displace_host) == 0, "vx_fused_row_features", "bad layout");
  VX_REQUIRE(C == 6 || C == 12, "vx_fused_row_features", "k0 channels must be 6 or 12");
  const VxGrid gs = make_grid(X, Y, Z, 1, 0, xyz_min_host, xyz_max_host);
  const VxGrid gk = make_grid(X, Y, Z, C, k0_channels_last, xyz_min_host, xyz_max_host);
  const VxPts pts{nullptr, ray_id, step_id, rays_start, rays_dir, stepdist};
  const int blocks = min(vx_blocks((int64_t)capacity * 4, 128), vx_num_sms() * 32);
  if (C == 6)
    k_row_features<6><<<blocks, 128, 0, st>>>(gs, sdf_grid, gk, k0_grid, pts, idx4, n_rows_dev, capacity, viewdirs, sdf_s,
                                              grad_s, voxel_size, use_grad_norm, lay, X1, X2);
  else
    k_row_features<12><<<blocks, 128, 0, st>>>(gs, sdf_grid, gk, k0_grid, pts, idx4, n_rows_dev, capacity, viewdirs, sdf_s,
                                               grad_s, voxel_size, use_grad_norm, lay, X1, X2);
  return vx_check_launch("vx_fused_row_features");
}

// rgb_logit.detach() into X2's last 3 feature columns (k_res, lib/voxurf_fine.py:741-744)
__global__ void k_fill_logit_cols(const float* __restrict__ logit, int ld_logit, const int* __restrict__ n_rows_dev,
                                  int capacity, int col, int ld2, float* __restrict__ X2) {
  const int n = min(*n_rows_dev, capacity);
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += gridDim.x * blockDim.x) {
    X2[(int64_t)row * ld2 + col] = logit[(int64_t)row * ld_logit];
    X2[(int64_t)row * ld2 + col + 1] = logit[(int64_t)row * ld_logit + 1];
    X2[(int64_t)row * ld2 + col + 2] = logit[(int64_t)row * ld_logit + 2];
  }
}

VX_API int vx_fused_fill_logit_cols(const float* logit, int ld_logit, const int* n_rows_dev, int capacity, int col, int ld2,
                                    float* X2, cudaStream_t st) {
  if (capacity <= 0) return 0;
  k_fill_logit_cols<<<min(vx_blocks(capacity, 256), vx_num_sms() * 8), 256, 0, st>>>(logit, ld_logit, n_rows_dev, capacity, col, ld2, X2);
  return vx_check_launch("vx_fused_fill_logit_cols");
}

// ---------------------------------------------------------------------------------------------
// S9: compositing + losses + their backward, one warp per ray (lib/voxurf_fine.py:749-763, run.py:604-636).
//   rgb = sigmoid(logit1); k_rgb = sigmoid(logit1.detach() + k_out)
//   rgb_marched0 = sum_i w_i rgb_i + alphainv_last * bg ; rgb_marched = clamp(sum_i w_i k_rgb_i + alphainv_last * bg, 0, 1)
//   loss = w_main * mse(rgb_marched, target) + w_rgb0 * mse(rgb_marched0, target) + w_ent * entropy(alphainv_last[N-1])
// train != 0: writes d_logit1, d_kout (rows), d_w (scattered to the M2 index of each row), d_last (rays), per-ray loss.
// Segmented sums are done by fixed-pattern warp reductions: deterministic, no atomics (torch_scatter uses atomics
// at segment boundaries).
// ---------------------------------------------------------------------------------------------
struct VxLossCfg {
  float w_main, w_rgb0, w_ent, ent_scale;  // ent_scale: 1 on a single GPU; see parallel.py for ray-sharded runs
  float inv_3n;                            // 1 / (3 * N_local): F.mse_loss mean
  float bg;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void k_composite_loss(const float* __restrict__ logit1, const float* __restrict__ k_out, int ld_out,
                                 const int* __restrict__ idx4, const int* __restrict__ off4, int capacity,
                                 const float* __restrict__ weight_s, const float* __restrict__ alphainv_last,
                                 const float* __restrict__ target, int n_rays, VxLossCfg cfg, int train,
                                 float* __restrict__ rgb_marched, float* __restrict__ rgb_marched0,
                                 float* __restrict__ d_logit1, float* __restrict__ d_kout, float* __restrict__ d_w_s,
                                 float* __restrict__ d_last, float* __restrict__ loss_ray) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * warps_per_block) {
    const int b = min(off4[r], capacity), e = min(off4[r + 1], capacity);
    float s0[3] = {0.f, 0.f, 0.f}, sk[3] = {0.f, 0.f, 0.f};
    for (int row = b + lane; row < e; row += 32) {
      const float w = weight_s[idx4[row]];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float l1 = logit1[(int64_t)row * ld_out + c];
        s0[c] += w * sigmoidf_(l1);
        sk[c] += w * sigmoidf_(l1 + k_out[(int64_t)row * ld_out + c]);
      }
    }
    const float al = alphainv_last[r];
    float g0[3], gk[3], lsum = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float m0 = warp_sum(s0[c]) + al * cfg.bg;
      const float mk_raw = warp_sum(sk[c]) + al * cfg.bg;
      const float mk = fminf(fmaxf(mk_raw, 0.f), 1.f);
      const float tg = target ? target[3 * r + c] : 0.f;
      if (lane == 0) { rgb_marched0[3 * r + c] = m0; rgb_marched[3 * r + c] = mk; }
      const float e0 = m0 - tg, ek = mk - tg;
      g0[c] = cfg.w_rgb0 * 2.f * e0 * cfg.inv_3n;
      gk[c] = (mk_raw >= 0.f && mk_raw <= 1.f) ? cfg.w_main * 2.f * ek * cfg.inv_3n : 0.f;
      lsum += cfg.w_rgb0 * e0 * e0 * cfg.inv_3n + cfg.w_main * ek * ek * cfg.inv_3n;
    }
    if (!train) continue;
    float dl = (g0[0] + g0[1] + g0[2] + gk[0] + gk[1] + gk[2]) * cfg.bg;
    if (r == n_rays - 1 && cfg.w_ent > 0.f && cfg.ent_scale != 0.f) {  // run.py:607-610: the last ray only
      const float pc = fminf(fmaxf(al, 1e-6f), 1.f - 1e-6f);
      lsum += cfg.ent_scale * cfg.w_ent * (-(pc * logf(pc) + (1.f - pc) * logf(1.f - pc)));
      if (al >= 1e-6f && al <= 1.f - 1e-6f) dl += cfg.ent_scale * cfg.w_ent * (logf(1.f - pc) - logf(pc));
    }
    if (lane == 0) { d_last[r] = dl; loss_ray[r] = lsum; }
    for (int row = b + lane; row < e; row += 32) {
      const int i = idx4[row];
      const float w = weight_s[i];
      float dw = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float l1 = logit1[(int64_t)row * ld_out + c];
        const float rgb = sigmoidf_(l1);
        const float krgb = sigmoidf_(l1 + k_out[(int64_t)row * ld_out + c]);
        dw += g0[c] * rgb + gk[c] * krgb;
        d_logit1[(int64_t)row * ld_out + c] = w * g0[c] * rgb * (1.f - rgb);
        d_kout[(int64_t)row * ld_out + c] = w * gk[c] * krgb * (1.f - krgb);
      }
      d_w_s[i] = dw;
    }
  }
}

__global__ void k_zero_tail_rows(float* __restrict__ a, float* __restrict__ b, int ld, const int* __restrict__ n_rows_dev,
                                 int capacity) {
  const int n = min(*n_rows_dev, capacity);
  for (int64_t t = (int64_t)n * ld + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (int64_t)capacity * ld;
       t += (int64_t)gridDim.x * blockDim.x) {
    a[t] = 0.f;
    b[t] = 0.f;
  }
}

VX_API int vx_fused_composite_loss(const float* logit1, const float* k_out, int ld_out, const int* idx4, const int* off4,
                                   int capacity, const float* weight_s, const float* alphainv_last, const float* target,
                                   int n_rays, float w_main, float w_rgb0, float w_ent, float ent_scale, float bg, int train,
                                   float* rgb_marched, float* rgb_marched0, float* d_logit1, float* d_kout, float* d_w_s,
                                   float* d_last, float* loss_ray, cudaStream_t st) {
  if (n_rays <= 0) return 0;
  VxLossCfg cfg{w_main, w_rgb0, w_ent, ent_scale, 1.f / (3.f * (float)n_rays), bg};
  if (train) {  // rows beyond M4 must carry zero output gradients (fixed-shape GEMM backward over the capacity)
    k_zero_tail_rows<<<vx_num_sms() * 2, 256, 0, st>>>(d_logit1, d_kout, ld_out, off4 + n_rays, capacity);
    int rc = vx_check_launch("vx_fused_composite_loss(tail)");
    if (rc) return rc;
  }
  const int blocks = min(vx_blocks((int64_t)n_rays * 32, 256), vx_num_sms() * 8);
  k_composite_loss<<<blocks, 256, 0, st>>>(logit1, k_out, ld_out, idx4, off4, capacity, weight_s, alphainv_last, target,
                                           n_rays, cfg, train, rgb_marched, rgb_marched0, d_logit1, d_kout, d_w_s, d_last,
                                           loss_ray);
  return vx_check_launch("vx_fused_composite_loss");
}

// render-only extras (lib/voxurf_fine.py:765-777): normal_marched (+1e-6 normalisation) and depth
__global__ void k_composite_aux(const int* __restrict__ idx4, const int* __restrict__ off4, int capacity,
                                const float* __restrict__ weight_s, const float* __restrict__ grad_s,
                                const int* __restrict__ step_id, float dist, int n_rays, float* __restrict__ normal_marched,
                                float* __restrict__ depth) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * warps_per_block) {
    const int b = min(off4[r], capacity), e = min(off4[r + 1], capacity);
    float nx = 0.f, ny = 0.f, nz = 0.f, dp = 0.f;
    for (int row = b + lane; row < e; row += 32) {
      const int i = idx4[row];
      const float w = weight_s[i];
      const float gx = grad_s[3 * i], gy = grad_s[3 * i + 1], gz = grad_s[3 * i + 2];
      const float nrm = sqrtf(gx * gx + gy * gy + gz * gz) + 1e-6f;
      nx += w * (gx / nrm); ny += w * (gy / nrm); nz += w * (gz / nrm);
      dp += __fmul_rn(__fmul_rn(w, (float)step_id[i]), dist);
    }
    nx = warp_sum(nx); ny = warp_sum(ny); nz = warp_sum(nz); dp = warp_sum(dp);
    if (lane == 0) {
      if (normal_marched) { normal_marched[3 * r] = nx; normal_marched[3 * r + 1] = ny; normal_marched[3 * r + 2] = nz; }
      if (depth) depth[r] = dp;
    }
  }
}

VX_API int vx_fused_composite_aux(const int* idx4, const int* off4, int capacity, const float* weight_s, const float* grad_s,
                                  const int* step_id, float dist, int n_rays, float* normal_marched, float* depth,
                                  cudaStream_t st) {
  if (n_rays <= 0) return 0;
  const int blocks = min(vx_blocks((int64_t)n_rays * 32, 256), vx_num_sms() * 8);
  k_composite_aux<<<blocks, 256, 0, st>>>(idx4, off4, capacity, weight_s, grad_s, step_id, dist, n_rays, normal_marched, depth);
  return vx_check_launch("vx_fused_composite_aux");
}

// ---------------------------------------------------------------------------------------------
// S11: backward of the row features: dX1 -> sdf-grid scatter (all_feat, all_grad) + d_sdf_s (centre sdf);
//      dX2 -> k0-grid scatter + d_grad_s (gradient feature of k_rgbnet).  One thread per row.
// ---------------------------------------------------------------------------------------------
// one corner of a channels-last k0 scatter: kC contiguous channels (vector atomics for fp32, scalar 64-bit adds otherwise)
template <int kC>
__device__ __forceinline__ void vx_k0_add(const VxAccF& a, int64_t base, const float* go, float w) {
  float* dst = a.p + base;
  if (kC % 4 == 0) {
#pragma unroll
    for (int c = 0; c < kC; c += 4)
      atomicAdd(reinterpret_cast<float4*>(dst + c), make_float4(go[c] * w, go[c + 1] * w, go[c + 2] * w, go[c + 3] * w));
  } else {
#pragma unroll
    for (int c = 0; c < kC; c += 2) atomicAdd(reinterpret_cast<float2*>(dst + c), make_float2(go[c] * w, go[c + 1] * w));
  }
}
template <int kC>
__device__ __forceinline__ void vx_k0_add(const VxAccQ& a, int64_t base, const float* go, float w) {
#pragma unroll
  for (int c = 0; c < kC; ++c) a.add(base + c, go[c] * w);
}

template <int kC, class AccS, class AccK>
__global__ void k_row_backward(VxGrid gs, const float* __restrict__ sdf_grid, VxGrid gk, VxPts pts,
                               const int* __restrict__ idx4, const int* __restrict__ n_rows_dev, int capacity,
                               float voxel_size, int use_grad_norm, VxRowLayout lay, const float* __restrict__ dX1,
                               const float* __restrict__ dX2, float* __restrict__ d_sdf_s, float* __restrict__ d_grad_s,
                               AccS sdf_grad, AccK k0_acc, bool has_k0, uint32_t* __restrict__ k0_touched) {
  const int n = min(*n_rows_dev, capacity);
  const int L = lay.L;
  const int col_sdf = 3 + 6 * lay.P + 3 + 6 * lay.Vp;
  const int col_grad2 = kC + 3 + 6 * lay.P2 + 3 + 6 * lay.V2;
  // four threads per row: sub-thread s scatters the displacements l = s, s+4, ...; sub-thread 0 also does the k0 scatter
  for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < n * 4; item += gridDim.x * blockDim.x) {
    const int row = item >> 2, sub = item & 3;
    const float* g1 = dX1 + (int64_t)row * lay.ld1;
    const float* g2 = dX2 + (int64_t)row * lay.ld2;
    const int i = idx4[row];
    float p[3];
    vx_load_pt(pts, i, p[0], p[1], p[2]);
    if (sub == 1) {
      d_sdf_s[i] = g1[col_sdf];
      d_grad_s[3 * i] = g2[col_grad2]; d_grad_s[3 * i + 1] = g2[col_grad2 + 1]; d_grad_s[3 * i + 2] = g2[col_grad2 + 2];
    }
    // ---- k0 scatter: corners 2 sub, 2 sub + 1 of the tap on each of the row's four threads
    {
      float go[kC];
      bool any = false;
#pragma unroll
      for (int c = 0; c < kC; ++c) { go[c] = g2[c]; any |= go[c] != 0.f; }
      if (any && has_k0) {
        float ix, iy, iz;
        point_to_index(gk, p[0], p[1], p[2], ix, iy, iz);
        VxTap t;
        vx_make_tap(ix, iy, iz, gk.X, gk.Y, gk.Z, t);
        const int64_t V = (int64_t)gk.X * gk.Y * gk.Z;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if ((k >> 1) != sub || t.off[k] < 0) continue;
          if (k0_touched) atomicOr(k0_touched + (t.off[k] >> 5), 1u << (t.off[k] & 31));
          if (gk.cl) {
            vx_k0_add<kC>(k0_acc, (int64_t)t.off[k] * kC, go, t.w[k]);
          } else {
#pragma unroll
            for (int c = 0; c < kC; ++c) k0_acc.add(c * V + t.off[k], go[c] * t.w[k]);
          }
        }
      }
    }
    // ---- sample_sdfs backward: the six taps of a displacement level are built once (axis parts) and used for the
    //      re-evaluation of the normalised gradient and for the scatter
    const float* dfeat = g1 + col_sdf + 1;
    const float* dgrad = dfeat + 6 * L;
    SdfTapCoords tc;
    sdf_tap_setup(gs, p[0], p[1], p[2], tc);
    const AxisTap sz = axis_tap(tc.c[0], gs.Z), sy = axis_tap(tc.c[1], gs.Y), sx = axis_tap(tc.c[2], gs.X);
    for (int l = sub; l < L; l += 4) {
      AxisTap dm[3], dp[3];
      float cm[3], cp[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        float ix, iy, iz;
        const int size = axis_size(gs, a);
        cm[a] = sdf_tap_coords(gs, tc, a, -lay.disp[l], ix, iy, iz);
        dm[a] = axis_tap(a == 0 ? ix : (a == 1 ? iy : iz), size);
        cp[a] = sdf_tap_coords(gs, tc, a, lay.disp[l], ix, iy, iz);
        dp[a] = axis_tap(a == 0 ? ix : (a == 1 ? iy : iz), size);
      }
      float dgr[3] = {dgrad[0 * L + l], dgrad[1 * L + l], dgrad[2 * L + l]};
      if (use_grad_norm && (dgr[0] != 0.f || dgr[1] != 0.f || dgr[2] != 0.f)) {
        float gr[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float fm = tap_eval_axes(sdf_grid, gs.Y, gs.Z, VX_AXES(a, dm[a], sz, sy, sx));
          const float fp = tap_eval_axes(sdf_grid, gs.Y, gs.Z, VX_AXES(a, dp[a], sz, sy, sx));
          gr[a] = __fdiv_rn(__fdiv_rn(__fsub_rn(fp, fm), __fsub_rn(cp[a], cm[a])), voxel_size);
        }
        const float nrm = sqrtf(gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2]);
        const float den = nrm + 1e-5f;
        const float dot = dgr[0] * gr[0] + dgr[1] * gr[1] + dgr[2] * gr[2];
        const float k = (nrm > 0.f) ? dot / (den * den * nrm) : 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) dgr[a] = dgr[a] / den - gr[a] * k;
      }
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        float dfm = dfeat[(a * 2 + 0) * L + l];
        float dfp = dfeat[(a * 2 + 1) * L + l];
        if (dgr[a] != 0.f) {
          const float d = (dgr[a] / voxel_size) / (cp[a] - cm[a]);
          dfp += d;
          dfm -= d;
        }
        tap_scatter_axes(sdf_grad, gs.Y, gs.Z, VX_AXES(a, dm[a], sz, sy, sx), dfm);
        tap_scatter_axes(sdf_grad, gs.Y, gs.Z, VX_AXES(a, dp[a], sz, sy, sx), dfp);
      }
    }
  }
}

template <class AccS, class AccK>
static int row_backward_impl(const float* sdf_grid, int X, int Y, int Z, int C, int k0_channels_last,
                             const float* xyz_min_host, const float* xyz_max_host, const int* ray_id, const int* step_id,
                             const float* rays_start, const float* rays_dir, float stepdist, const int* idx4,
                             const int* n_rows_dev, int capacity, float voxel_size, int use_grad_norm, int P, int Vp,
                             int P2, int V2, const float* displace_host, int L, int ld1, int ld2, const float* dX1,
                             const float* dX2, float* d_sdf_s, float* d_grad_s, AccS sdf_acc, AccK k0_acc, bool has_k0,
                             uint32_t* k0_touched, cudaStream_t st) {
  if (capacity <= 0) return 0;
  VxRowLayout lay;
  VX_REQUIRE(fill_layout(lay, P, Vp, P2, V2, L, C, ld1, ld2, displace_host) == 0, "vx_fused_row_backward", "bad layout");
  VX_REQUIRE(C == 6 || C == 12, "vx_fused_row_backward", "k0 channels must be 6 or 12");
  const VxGrid gs = make_grid(X, Y, Z, 1, 0, xyz_min_host, xyz_max_host);
  const VxGrid gk = make_grid(X, Y, Z, C, k0_channels_last, xyz_min_host, xyz_max_host);
  const VxPts pts{nullptr, ray_id, step_id, rays_start, rays_dir, stepdist};
  const int blocks = min(vx_blocks((int64_t)capacity * 4, 128), vx_num_sms() * 32);
  if (C == 6)
    k_row_backward<6, AccS, AccK><<<blocks, 128, 0, st>>>(gs, sdf_grid, gk, pts, idx4, n_rows_dev, capacity, voxel_size, use_grad_norm,
                                                          lay, dX1, dX2, d_sdf_s, d_grad_s, sdf_acc, k0_acc, has_k0, k0_touched);
  else
    k_row_backward<12, AccS, AccK><<<blocks, 128, 0, st>>>(gs, sdf_grid, gk, pts, idx4, n_rows_dev, capacity, voxel_size, use_grad_norm,
                                                           lay, dX1, dX2, d_sdf_s, d_grad_s, sdf_acc, k0_acc, has_k0, k0_touched);
  return vx_check_launch("vx_fused_row_backward");
}

VX_API int vx_fused_row_backward(const float* sdf_grid, int X, int Y, int Z, int C, int k0_channels_last,
                                 const float* xyz_min_host, const float* xyz_max_host, const int* ray_id, const int* step_id,
                                 const float* rays_start, const float* rays_dir, float stepdist, const int* idx4,
                                 const int* n_rows_dev, int capacity, float voxel_size, int use_grad_norm, int P, int Vp,
                                 int P2, int V2, const float* displace_host, int L, int ld1, int ld2, const float* dX1,
                                 const float* dX2, float* d_sdf_s, float* d_grad_s, float* sdf_grad, float* k0_grad,
                                 uint32_t* k0_touched, cudaStream_t st) {
  return row_backward_impl(sdf_grid, X, Y, Z, C, k0_channels_last, xyz_min_host, xyz_max_host, ray_id, step_id, rays_start, rays_dir,
                           stepdist, idx4, n_rows_dev, capacity, voxel_size, use_grad_norm, P, Vp, P2, V2, displace_host, L, ld1, ld2,
                           dX1, dX2, d_sdf_s, d_grad_s, VxAccF{sdf_grad}, VxAccF{k0_grad}, k0_grad != nullptr, k0_touched, st);
}

// the same scatter into 64-bit fixed-point accumulators (order-independent sums: bit-reproducible gradients); acc_scale =
// accumulator units per 1.0, e.g. 2^52.  vx_fx_accumulate folds the accumulators into the fp32 gradient grids.
VX_API int vx_fused_row_backward_fx(const float* sdf_grid, int X, int Y, int Z, int C, int k0_channels_last,
                                    const float* xyz_min_host, const float* xyz_max_host, const int* ray_id, const int* step_id,
                                    const float* rays_start, const float* rays_dir, float stepdist, const int* idx4,
                                    const int* n_rows_dev, int capacity, float voxel_size, int use_grad_norm, int P, int Vp,
                                    int P2, int V2, const float* displace_host, int L, int ld1, int ld2, const float* dX1,
                                    const float* dX2, float* d_sdf_s, float* d_grad_s, int64_t* sdf_acc, int64_t* k0_acc,
                                    float acc_scale, uint32_t* k0_touched, cudaStream_t st) {
  VX_REQUIRE(sdf_acc != nullptr && acc_scale > 0.f, "vx_fused_row_backward_fx", "sdf_acc / acc_scale required");
  return row_backward_impl(sdf_grid, X, Y, Z, C, k0_channels_last, xyz_min_host, xyz_max_host, ray_id, step_id, rays_start, rays_dir,
                           stepdist, idx4, n_rows_dev, capacity, voxel_size, use_grad_norm, P, Vp, P2, V2, displace_host, L, ld1, ld2,
                           dX1, dX2, d_sdf_s, d_grad_s, VxAccQ{reinterpret_cast<unsigned long long*>(sdf_acc), (double)acc_scale},
                           VxAccQ{reinterpret_cast<unsigned long long*>(k0_acc), (double)acc_scale}, k0_acc != nullptr, k0_touched, st);
}

// ---------------------------------------------------------------------------------------------
// S13: NeuS-alpha backward + the M2-level scatter into the sdf grid (7 taps per sample).
// d_alpha comes from vx_alpha2weight_seg_backward (zero for dropped samples and behind the early exit);
// d_sdf_s / d_grad_s carry the MLP-feature gradients of the rows.  Samples whose gradients are all exactly zero
// (the vast majority late in training) are skipped before any tap is recomputed.
// ---------------------------------------------------------------------------------------------
template <class Acc>
__global__ void k_alpha_sdf_bwd(VxGrid g, VxPts pts, const int* __restrict__ n_dev, const float* __restrict__ viewdirs,
                                const float* __restrict__ sdf, const float* __restrict__ grad,
                                const uint8_t* __restrict__ keep, const float* __restrict__ d_alpha,
                                const float* __restrict__ d_sdf_s, const float* __restrict__ d_grad_s, float voxel_size,
                                float dist, float inv_s, Acc sdf_grad, const float* __restrict__ inv_s_dev) {
  const int64_t n = *n_dev;
  if (inv_s_dev) inv_s = __ldg(inv_s_dev);
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    if (!keep[p]) continue;
    const float ga = d_alpha[p];
    float ds = d_sdf_s[p];
    float dg[3] = {d_grad_s[3 * p], d_grad_s[3 * p + 1], d_grad_s[3 * p + 2]};  // x,y,z
    if (ga != 0.f) {
      const int r = pts.ray_id[p];
      const float vx = viewdirs[3 * r], vy = viewdirs[3 * r + 1], vz = viewdirs[3 * r + 2];
      const float gx = grad[3 * p], gy = grad[3 * p + 1], gz = grad[3 * p + 2];
      const float s = sdf[p];
      const float true_cos = __fadd_rn(__fadd_rn(__fmul_rn(vx, gx), __fmul_rn(vy, gy)), __fmul_rn(vz, gz));
      const float iter_cos = -fmaxf(-true_cos, 0.f);
      const float h = __fmul_rn(__fmul_rn(iter_cos, dist), 0.5f);
      const float prev = sigmoidf_(__fmul_rn(__fsub_rn(s, h), inv_s));
      const float next = sigmoidf_(__fmul_rn(__fadd_rn(s, h), inv_s));
      const float u = __fadd_rn(__fsub_rn(prev, next), 1e-5f);
      const float c = __fadd_rn(prev, 1e-5f);
      const float a = __fdiv_rn(u, c);
      if (a >= 0.f && a <= 1.f) {
        const float du = ga / c;
        const float dc = -ga * u / (c * c);
        const float dep = (du + dc) * prev * (1.f - prev) * inv_s;
        const float den = (-du) * next * (1.f - next) * inv_s;
        ds += dep + den;
        const float dcos = (true_cos < 0.f) ? (den - dep) * dist * 0.5f : 0.f;
        dg[0] += dcos * vx; dg[1] += dcos * vy; dg[2] += dcos * vz;
      }
    }
    if (ds == 0.f && dg[0] == 0.f && dg[1] == 0.f && dg[2] == 0.f) continue;
    float px, py, pz;
    vx_load_pt(pts, p, px, py, pz);
    VxTap t;
    float ix, iy, iz;
    if (ds != 0.f) {
      point_to_index(g, px, py, pz, ix, iy, iz);
      vx_make_tap(ix, iy, iz, g.X, g.Y, g.Z, t);
      vx_tap_scatter(sdf_grad, t, ds);
    }
    SdfTapCoords tc;
    sdf_tap_setup(g, px, py, pz, tc);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float dga = dg[2 - a];  // reference axis a = z,y,x  <->  xyz component 2 - a
      if (dga == 0.f) continue;
      const float cm = sdf_tap_coords(g, tc, a, -1.f, ix, iy, iz);
      vx_make_tap(ix, iy, iz, g.X, g.Y, g.Z, t);
      VxTap tp;
      const float cp = sdf_tap_coords(g, tc, a, 1.f, ix, iy, iz);
      vx_make_tap(ix, iy, iz, g.X, g.Y, g.Z, tp);
      const float d = (dga / voxel_size) / (cp - cm);
      vx_tap_scatter(sdf_grad, t, -d);
      vx_tap_scatter(sdf_grad, tp, d);
    }
  }
}

VX_API int vx_fused_alpha_sdf_backward(int X, int Y, int Z, const float* xyz_min_host, const float* xyz_max_host,
                                       const int* ray_id, const int* step_id, const float* rays_start, const float* rays_dir,
                                       float stepdist, const int* n_dev, const float* viewdirs, const float* sdf,
                                       const float* grad, const uint8_t* keep, const float* d_alpha, const float* d_sdf_s,
                                       const float* d_grad_s, float voxel_size, float dist, float inv_s, float* sdf_grad,
                                       const float* inv_s_dev, cudaStream_t st) {
  VX_REQUIRE(n_dev != nullptr, "vx_fused_alpha_sdf_backward", "n_dev required");
  const VxGrid g = make_grid(X, Y, Z, 1, 0, xyz_min_host, xyz_max_host);
  const VxPts pts{nullptr, ray_id, step_id, rays_start, rays_dir, stepdist};
  k_alpha_sdf_bwd<VxAccF><<<vx_num_sms() * 8, 256, 0, st>>>(g, pts, n_dev, viewdirs, sdf, grad, keep, d_alpha, d_sdf_s, d_grad_s,
                                                            voxel_size, dist, inv_s, VxAccF{sdf_grad}, inv_s_dev);
  return vx_check_launch("vx_fused_alpha_sdf_backward");
}

VX_API int vx_fused_alpha_sdf_backward_fx(int X, int Y, int Z, const float* xyz_min_host, const float* xyz_max_host,
                                          const int* ray_id, const int* step_id, const float* rays_start, const float* rays_dir,
                                          float stepdist, const int* n_dev, const float* viewdirs, const float* sdf,
                                          const float* grad, const uint8_t* keep, const float* d_alpha, const float* d_sdf_s,
                                          const float* d_grad_s, float voxel_size, float dist, float inv_s, int64_t* sdf_acc,
                                          float acc_scale, const float* inv_s_dev, cudaStream_t st) {
  VX_REQUIRE(n_dev != nullptr && sdf_acc != nullptr && acc_scale > 0.f, "vx_fused_alpha_sdf_backward_fx", "n_dev / sdf_acc / acc_scale required");
  const VxGrid g = make_grid(X, Y, Z, 1, 0, xyz_min_host, xyz_max_host);
  const VxPts pts{nullptr, ray_id, step_id, rays_start, rays_dir, stepdist};
  k_alpha_sdf_bwd<VxAccQ><<<vx_num_sms() * 8, 256, 0, st>>>(g, pts, n_dev, viewdirs, sdf, grad, keep, d_alpha, d_sdf_s, d_grad_s, voxel_size,
                                                            dist, inv_s, VxAccQ{reinterpret_cast<unsigned long long*>(sdf_acc), (double)acc_scale},
                                                            inv_s_dev);
  return vx_check_launch("vx_fused_alpha_sdf_backward_fx");
}

// grad[i] += acc[i] / scale, acc[i] = 0 where acc[i] != 0 (dense read of the accumulators, sparse writes).  touched
// (optional, one bit per group of `group` consecutive elements): only the flagged groups are visited.
__global__ void k_fx_accumulate(long long* __restrict__ acc, int64_t n, double inv_scale, float* __restrict__ grad,
                                const uint32_t* __restrict__ touched, int group) {
  if (touched) {
    const int64_t n_vox = n / group, n_words = (n_vox + 31) >> 5;
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int64_t w = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); w < n_words; w += (int64_t)gridDim.x * wpb) {
      const uint32_t bits = touched[w];
      if (bits == 0u) continue;
      for (int64_t e = lane; e < (int64_t)32 * group; e += 32) {
        const int64_t i = w * 32 * group + e;
        if (i >= n || !((bits >> (e / group)) & 1u)) continue;
        const long long a = acc[i];
        if (a != 0) { grad[i] += (float)((double)a * inv_scale); acc[i] = 0; }
      }
    }
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const long long a = acc[i];
    if (a != 0) { grad[i] += (float)((double)a * inv_scale); acc[i] = 0; }
  }
}

VX_API int vx_fx_accumulate(int64_t* acc, int64_t n, float acc_scale, float* grad, const uint32_t* touched, int group,
                            cudaStream_t st) {
  if (n <= 0) return 0;
  VX_REQUIRE(acc && grad && acc_scale > 0.f && group >= 1, "vx_fx_accumulate", "bad arguments");
  const int64_t work = touched ? ((n / group + 31) / 32) * 32 : n;
  k_fx_accumulate<<<(int)min((int64_t)vx_blocks(work, 256), (int64_t)vx_num_sms() * 16), 256, 0, st>>>(
      reinterpret_cast<long long*>(acc), n, 1.0 / (double)acc_scale, grad, touched, group);
  return vx_check_launch("vx_fx_accumulate");
}


// ---------------------------------------------------------------------------------------------
// Ray-sharded data parallelism: instead of all-reducing the dense k0 gradient grid (0.8 GB at 256^3 x 12), every
// rank exports its MLP rows' k0 feature gradients -- (xyz, dk0[C]) per row, ~3 MB -- the ranks all-gather them and each
// replays the trilinear scatter of ALL ranks' rows into its own gradient grid (vx_grid_gather_backward).
// `scale` = 1 / world (the global-batch gradient is the mean of the per-rank gradients).
// ---------------------------------------------------------------------------------------------
__global__ void k_export_k0_rows(VxPts pts, const int* __restrict__ idx4, const int* __restrict__ n_rows_dev, int capacity,
                                 const float* __restrict__ dX2, int ld2, int C, float scale, float* __restrict__ xyz_out,
                                 float* __restrict__ g_out, int* __restrict__ n_out) {
  const int n = min(*n_rows_dev, capacity);
  if (n_out && blockIdx.x == 0 && threadIdx.x == 0) *n_out = n;   // travels with the rows: receivers scatter n rows, not `capacity`
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < capacity; row += gridDim.x * blockDim.x) {
    float p[3] = {0.f, 0.f, 0.f};
    if (row < n) vx_load_pt(pts, idx4[row], p[0], p[1], p[2]);
    xyz_out[3 * row] = p[0]; xyz_out[3 * row + 1] = p[1]; xyz_out[3 * row + 2] = p[2];
    for (int c = 0; c < C; ++c) g_out[(int64_t)row * C + c] = (row < n) ? dX2[(int64_t)row * ld2 + c] * scale : 0.f;
  }
}

VX_API int vx_fused_export_k0_rows(const int* ray_id, const int* step_id, const float* rays_start, const float* rays_dir,
                                   float stepdist, const int* idx4, const int* n_rows_dev, int capacity, const float* dX2,
                                   int ld2, int C, float scale, float* xyz_out, float* g_out, int* n_out, cudaStream_t st) {
  if (capacity <= 0) return 0;
  const VxPts pts{nullptr, ray_id, step_id, rays_start, rays_dir, stepdist};
  k_export_k0_rows<<<min(vx_blocks(capacity, 256), vx_num_sms() * 8), 256, 0, st>>>(pts, idx4, n_rows_dev, capacity, dX2, ld2, C,
                                                                                    scale, xyz_out, g_out, n_out);
  return vx_check_launch("vx_fused_export_k0_rows");
}

// small helpers for the host orchestration ----------------------------------------------------
__global__ void k_sum_f32(const float* __restrict__ x, int n, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) out[0] = t;
  }
}

VX_API int vx_sum_f32(const float* x, int n, float* out, cudaStream_t st) {
  k_sum_f32<<<1, 1024, 0, st>>>(x, n, out);
  return vx_check_launch("vx_sum_f32");
}
