// Ray / AABB sampling (SURVEY.md 8a rows A1-A6').
//
// Legacy surface: one entry point per reference pybind function of
// /root/reference/lib/cuda/render_utils.cpp:170-184, with the arithmetic spelled in the same
// expression order so that nvcc's default FMA contraction produces bit-identical t_min/t_max,
// N_steps, ray_id, step_id, points and masks (render_utils_kernel.cu:12-242, 367-424).
//
// Fused surface: vx_ray_march_* replaces sample_pts_on_rays + the in-bbox compaction
// (voxurf_fine.py:612-616) + MaskCache.forward + its compaction (voxurf_fine.py:631-636) by
// one warp-per-ray pass that ballots the keep-flags into per-ray bit words and one pass that
// expands the bits into the compact, ray-sorted (ray_id, step_id) list.  Nothing of size M0
// (~5 M samples, 145 MB in the reference) is ever written.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// per-ray quantities (render_utils_kernel.cu:12-79)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ray_t_minmax(const float* __restrict__ o, const float* __restrict__ d,
                                             const float* __restrict__ xyz_min, const float* __restrict__ xyz_max,
                                             float near, float far, float& t_min, float& t_max) {
  float vx = ((d[0] == 0) ? 1e-6 : d[0]);
  float vy = ((d[1] == 0) ? 1e-6 : d[1]);
  float vz = ((d[2] == 0) ? 1e-6 : d[2]);
  float ax = (xyz_max[0] - o[0]) / vx;
  float ay = (xyz_max[1] - o[1]) / vy;
  float az = (xyz_max[2] - o[2]) / vz;
  float bx = (xyz_min[0] - o[0]) / vx;
  float by = (xyz_min[1] - o[1]) / vy;
  float bz = (xyz_min[2] - o[2]) / vz;
  t_min = max(min(max(max(min(ax, bx), min(ay, by)), min(az, bz)), far), near);
  t_max = max(min(min(min(max(ax, bx), max(ay, by)), max(az, bz)), far), near);
}

__device__ __forceinline__ float ray_norm(const float* __restrict__ d) {
  return sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
}

__device__ __forceinline__ int64_t ray_n_samples(float t_min, float t_max, float rnorm, float stepdist) {
  return max(ceil((t_max - t_min) * rnorm / stepdist), 1.);
}

__global__ void k_infer_t_minmax(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                 const float* __restrict__ xyz_min, const float* __restrict__ xyz_max, float near,
                                 float far, int n_rays, float* __restrict__ t_min, float* __restrict__ t_max) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rays) ray_t_minmax(rays_o + 3 * r, rays_d + 3 * r, xyz_min, xyz_max, near, far, t_min[r], t_max[r]);
}

__global__ void k_infer_n_samples(const float* __restrict__ rays_d, const float* __restrict__ t_min,
                                  const float* __restrict__ t_max, float stepdist, int n_rays,
                                  int64_t* __restrict__ n_samples) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rays) n_samples[r] = ray_n_samples(t_min[r], t_max[r], ray_norm(rays_d + 3 * r), stepdist);
}

__global__ void k_infer_ray_start_dir(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                      const float* __restrict__ t_min, int n_rays, float* __restrict__ rays_start,
                                      float* __restrict__ rays_dir) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rays) {
    const int o = 3 * r;
    const float rnorm = ray_norm(rays_d + o);
    rays_start[o] = rays_o[o] + rays_d[o] * t_min[r];
    rays_start[o + 1] = rays_o[o + 1] + rays_d[o + 1] * t_min[r];
    rays_start[o + 2] = rays_o[o + 2] + rays_d[o + 2] * t_min[r];
    rays_dir[o] = rays_d[o] / rnorm;
    rays_dir[o + 1] = rays_d[o + 1] / rnorm;
    rays_dir[o + 2] = rays_d[o + 2] / rnorm;
  }
}

VX_API int vx_infer_t_minmax(const float* rays_o, const float* rays_d, const float* xyz_min, const float* xyz_max,
                             float near, float far, int n_rays, float* t_min, float* t_max, cudaStream_t st) {
  if (n_rays <= 0) return 0;
  k_infer_t_minmax<<<vx_blocks(n_rays, 256), 256, 0, st>>>(rays_o, rays_d, xyz_min, xyz_max, near, far, n_rays, t_min, t_max);
  return vx_check_launch("vx_infer_t_minmax");
}

VX_API int vx_infer_n_samples(const float* rays_d, const float* t_min, const float* t_max, float stepdist, int n_rays,
                              int64_t* n_samples, cudaStream_t st) {
  if (n_rays <= 0) return 0;
  k_infer_n_samples<<<vx_blocks(n_rays, 256), 256, 0, st>>>(rays_d, t_min, t_max, stepdist, n_rays, n_samples);
  return vx_check_launch("vx_infer_n_samples");
}

VX_API int vx_infer_ray_start_dir(const float* rays_o, const float* rays_d, const float* t_min, int n_rays,
                                  float* rays_start, float* rays_dir, cudaStream_t st) {
  if (n_rays <= 0) return 0;
  k_infer_ray_start_dir<<<vx_blocks(n_rays, 256), 256, 0, st>>>(rays_o, rays_d, t_min, n_rays, rays_start, rays_dir);
  return vx_check_launch("vx_infer_ray_start_dir");
}

// ---------------------------------------------------------------------------------------------
// ray setup: everything per-ray in one launch spread over the SMs, then a one-CTA exclusive scan of
// N_steps (n_rays is a batch of ~8192; the reference uses torch cumsum + sum().item(),
// render_utils_kernel.cu:210-212).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_ray_setup(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                   const float* __restrict__ xyz_min, const float* __restrict__ xyz_max,
                                                   float near, float far, float stepdist, int n_rays,
                                                   float* __restrict__ t_min, float* __restrict__ t_max,
                                                   int64_t* __restrict__ n_steps, float* __restrict__ rays_start,
                                                   float* __restrict__ rays_dir) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  const int o = 3 * r;
  float tmn, tmx;
  ray_t_minmax(rays_o + o, rays_d + o, xyz_min, xyz_max, near, far, tmn, tmx);
  const float rnorm = ray_norm(rays_d + o);
  t_min[r] = tmn; t_max[r] = tmx; n_steps[r] = ray_n_samples(tmn, tmx, rnorm, stepdist);
  rays_start[o] = rays_o[o] + rays_d[o] * tmn;
  rays_start[o + 1] = rays_o[o + 1] + rays_d[o + 1] * tmn;
  rays_start[o + 2] = rays_o[o + 2] + rays_d[o + 2] * tmn;
  rays_dir[o] = rays_d[o] / rnorm;
  rays_dir[o + 1] = rays_d[o + 1] / rnorm;
  rays_dir[o + 2] = rays_d[o + 2] / rnorm;
}

// exclusive scan of n_steps[0..n) -> offsets[0..n], one CTA: every thread owns 8 consecutive entries per pass
constexpr int kScanPerThread = 8;
__global__ void __launch_bounds__(1024) k_ray_offsets(const int64_t* __restrict__ n_steps, int n_rays,
                                                      int64_t* __restrict__ offsets) {
  __shared__ int64_t warp_sum[32];
  __shared__ int64_t carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n_rays; base += blockDim.x * kScanPerThread) {
    const int r0 = base + threadIdx.x * kScanPerThread;
    int64_t n[kScanPerThread];
    int64_t tot = 0;
#pragma unroll
    for (int j = 0; j < kScanPerThread; ++j) {
      n[j] = (r0 + j < n_rays) ? n_steps[r0 + j] : 0;
      tot += n[j];
    }
    int64_t x = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int64_t w = warp_sum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int64_t y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += y;
      }
      warp_sum[lane] = w;
    }
    __syncthreads();
    const int64_t incl = x + (warp > 0 ? warp_sum[warp - 1] : 0) + carry_s;
    int64_t run = incl - tot;
#pragma unroll
    for (int j = 0; j < kScanPerThread; ++j) {
      if (r0 + j < n_rays) offsets[r0 + j] = run;
      run += n[j];
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n_rays] = carry_s;
}

VX_API int vx_ray_setup(const float* rays_o, const float* rays_d, const float* xyz_min, const float* xyz_max, float near,
                        float far, float stepdist, int n_rays, float* t_min, float* t_max, int64_t* n_steps,
                        float* rays_start, float* rays_dir, int64_t* offsets, cudaStream_t st) {
  VX_REQUIRE(n_rays >= 0, "vx_ray_setup", "n_rays < 0");
  if (n_rays > 0)
    k_ray_setup<<<vx_blocks(n_rays, 128), 128, 0, st>>>(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist, n_rays, t_min,
                                                       t_max, n_steps, rays_start, rays_dir);
  k_ray_offsets<<<1, 1024, 0, st>>>(n_steps, n_rays, offsets);
  return vx_check_launch("vx_ray_setup");
}

// ---------------------------------------------------------------------------------------------
// legacy flat fill (render_utils_kernel.cu:144-194): one warp per ray writes its N_steps[r]
// consecutive slots -- coalesced stores, no scatter-1 + cumsum + per-sample ray lookup.
// ---------------------------------------------------------------------------------------------
__global__ void k_sample_fill(const float* __restrict__ rays_start, const float* __restrict__ rays_dir,
                              const float* __restrict__ xyz_min, const float* __restrict__ xyz_max,
                              const int64_t* __restrict__ offsets, int n_rays, float stepdist,
                              float* __restrict__ rays_pts, bool* __restrict__ mask_outbbox,
                              int64_t* __restrict__ ray_id, int64_t* __restrict__ step_id) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const float mnx = xyz_min[0], mny = xyz_min[1], mnz = xyz_min[2];
  const float mxx = xyz_max[0], mxy = xyz_max[1], mxz = xyz_max[2];
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * warps_per_block) {
    const int64_t beg = offsets[r];
    const int n = (int)(offsets[r + 1] - beg);
    const float sx = rays_start[3 * r], sy = rays_start[3 * r + 1], sz = rays_start[3 * r + 2];
    const float dx = rays_dir[3 * r], dy = rays_dir[3 * r + 1], dz = rays_dir[3 * r + 2];
    for (int s = lane; s < n; s += 32) {
      const int64_t idx = beg + s;
      const float dist = stepdist * s;
      const float px = sx + dx * dist;
      const float py = sy + dy * dist;
      const float pz = sz + dz * dist;
      rays_pts[3 * idx] = px; rays_pts[3 * idx + 1] = py; rays_pts[3 * idx + 2] = pz;
      mask_outbbox[idx] = (mnx > px) | (mny > py) | (mnz > pz) | (mxx < px) | (mxy < py) | (mxz < pz);
      ray_id[idx] = r;
      step_id[idx] = s;
    }
  }
}

VX_API int vx_sample_fill(const float* rays_start, const float* rays_dir, const float* xyz_min, const float* xyz_max,
                          const int64_t* offsets, int n_rays, float stepdist, float* rays_pts, bool* mask_outbbox,
                          int64_t* ray_id, int64_t* step_id, cudaStream_t st) {
  if (n_rays <= 0) return 0;
  const int blocks = min(vx_blocks((int64_t)n_rays * 32, 256), vx_num_sms() * 8);
  k_sample_fill<<<blocks, 256, 0, st>>>(rays_start, rays_dir, xyz_min, xyz_max, offsets, n_rays, stepdist, rays_pts,
                                        mask_outbbox, ray_id, step_id);
  return vx_check_launch("vx_sample_fill");
}

// render_utils_kernel.cu:245-293 (fixed-count NDC sampling) and :301-360 (inverted-sphere background)
__global__ void k_sample_ndc(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                             const float* __restrict__ xyz_min, const float* __restrict__ xyz_max, int N_samples,
                             int n_rays, float* __restrict__ rays_pts, bool* __restrict__ mask_outbbox) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < (int64_t)N_samples * n_rays) {
    const int i_ray = idx / N_samples;
    const int i_step = idx % N_samples;
    const int offset_r = i_ray * 3;
    const float dist = ((float)i_step) / (N_samples - 1);
    const float px = rays_o[offset_r] + rays_d[offset_r] * dist;
    const float py = rays_o[offset_r + 1] + rays_d[offset_r + 1] * dist;
    const float pz = rays_o[offset_r + 2] + rays_d[offset_r + 2] * dist;
    rays_pts[idx * 3] = px; rays_pts[idx * 3 + 1] = py; rays_pts[idx * 3 + 2] = pz;
    mask_outbbox[idx] = (xyz_min[0] > px) | (xyz_min[1] > py) | (xyz_min[2] > pz) | (xyz_max[0] < px) |
                        (xyz_max[1] < py) | (xyz_max[2] < pz);
  }
}

__global__ void k_sample_bg(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                            const float* __restrict__ t_max, float bg_preserve, int N_samples, int n_rays,
                            float* __restrict__ rays_pts) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < (int64_t)N_samples * n_rays) {
    const int i_ray = idx / N_samples;
    const int i_step = idx % N_samples;
    const int offset_r = i_ray * 3;
    const float t_inner = t_max[i_ray];
    const float ori_t_outer = t_inner - 1. + 1. / (1. - ((float)i_step) / N_samples);
    const float x = rays_o[offset_r] + rays_d[offset_r] * ori_t_outer;
    const float y = rays_o[offset_r + 1] + rays_d[offset_r + 1] * ori_t_outer;
    const float z = rays_o[offset_r + 2] + rays_d[offset_r + 2] * ori_t_outer;
    const float t_outer = sqrt(x * x + y * y + z * z);
    const float m = max(abs(x), max(abs(y), abs(z)));
    const float R_outer = t_outer / m;
    const float o2i_p = R_outer * R_outer / (t_outer * t_outer) * (1. - bg_preserve) + R_outer / t_outer * bg_preserve;
    rays_pts[idx * 3] = x * o2i_p; rays_pts[idx * 3 + 1] = y * o2i_p; rays_pts[idx * 3 + 2] = z * o2i_p;
  }
}

VX_API int vx_sample_ndc_pts_on_rays(const float* rays_o, const float* rays_d, const float* xyz_min, const float* xyz_max,
                                     int n_samples, int n_rays, float* rays_pts, bool* mask_outbbox, cudaStream_t st) {
  const int64_t n = (int64_t)n_samples * n_rays;
  if (n <= 0) return 0;
  k_sample_ndc<<<vx_blocks(n, 256), 256, 0, st>>>(rays_o, rays_d, xyz_min, xyz_max, n_samples, n_rays, rays_pts, mask_outbbox);
  return vx_check_launch("vx_sample_ndc_pts_on_rays");
}

VX_API int vx_sample_bg_pts_on_rays(const float* rays_o, const float* rays_d, const float* t_max, float bg_preserve,
                                    int n_samples, int n_rays, float* rays_pts, cudaStream_t st) {
  const int64_t n = (int64_t)n_samples * n_rays;
  if (n <= 0) return 0;
  k_sample_bg<<<vx_blocks(n, 256), 256, 0, st>>>(rays_o, rays_d, t_max, bg_preserve, n_samples, n_rays, rays_pts);
  return vx_check_launch("vx_sample_bg_pts_on_rays");
}

// ---------------------------------------------------------------------------------------------
// bool-grid free-space lookup (render_utils_kernel.cu:367-424)
// ---------------------------------------------------------------------------------------------
__global__ void k_maskcache_lookup(const bool* __restrict__ world, const float* __restrict__ xyz, bool* __restrict__ out,
                                   const float* __restrict__ scale, const float* __restrict__ shift, int sz_i, int sz_j,
                                   int sz_k, int64_t n_pts) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_pts) {
    const int i = round(xyz[3 * p] * scale[0] + shift[0]);
    const int j = round(xyz[3 * p + 1] * scale[1] + shift[1]);
    const int k = round(xyz[3 * p + 2] * scale[2] + shift[2]);
    bool v = false;
    if ((0 <= i) && (i < sz_i) && (0 <= j) && (j < sz_j) && (0 <= k) && (k < sz_k))
      v = world[(int64_t)i * sz_j * sz_k + (int64_t)j * sz_k + k];
    out[p] = v;
  }
}

VX_API int vx_maskcache_lookup(const bool* world, const float* xyz, const float* scale, const float* shift, int sz_i,
                               int sz_j, int sz_k, int64_t n_pts, bool* out, cudaStream_t st) {
  if (n_pts <= 0) return 0;
  k_maskcache_lookup<<<vx_blocks(n_pts, 256), 256, 0, st>>>(world, xyz, out, scale, shift, sz_i, sz_j, sz_k, n_pts);
  return vx_check_launch("vx_maskcache_lookup");
}

// ---------------------------------------------------------------------------------------------
// Fused march: keep-flags -> bit words -> compact list.
// keep(sample) = in-bbox (render_utils_kernel.cu:191-192 negated)
//              & [mask cache]  1 - exp(-softplus(trilinear(density) + act_shift) * ratio) >= thres
//                              (voxurf_fine.py:930-942; F.softplus: x > 20 ? x : log1p(exp(x)))
// ---------------------------------------------------------------------------------------------
struct VxMaskCache {
  const float* density;  // (X,Y,Z) max-pooled coarse density, or nullptr = no mask cache
  const uint8_t* cells;  // optional (X,Y,Z) per-cell verdicts from vx_mask_cache_cells, or nullptr
  int X, Y, Z;
  float min[3], max[3];
  float act_shift, voxel_size_ratio, thres;
};

__device__ __forceinline__ float mask_cache_alpha(const VxMaskCache& mc, float d) {
  const float x = d + mc.act_shift;
  const float sp = (x > 20.f) ? x : log1pf(expf(x));
  return 1.f - expf(__fmul_rn(-sp, mc.voxel_size_ratio));
}

__device__ __forceinline__ bool mask_cache_keep(const VxMaskCache& mc, float px, float py, float pz) {
  const float iz = vx_unnorm_coord(vx_norm_coord(px, mc.min[0], mc.max[0]), mc.X);
  const float iy = vx_unnorm_coord(vx_norm_coord(py, mc.min[1], mc.max[1]), mc.Y);
  const float ix = vx_unnorm_coord(vx_norm_coord(pz, mc.min[2], mc.max[2]), mc.Z);
  if (mc.cells) {
    // a sample strictly inside the lattice (all 8 corners valid; NaNs fail the comparisons) whose cell has a verdict
    if (ix >= 0.f && ix < (float)(mc.Z - 1) && iy >= 0.f && iy < (float)(mc.Y - 1) && iz >= 0.f && iz < (float)(mc.X - 1)) {
      const uint8_t c = __ldg(mc.cells + ((int)iz * mc.Y + (int)iy) * mc.Z + (int)ix);
      if (c < 2) return c != 0;
    }
  }
  VxTap t;
  vx_make_tap(ix, iy, iz, mc.X, mc.Y, mc.Z, t);
  const float d = vx_tap_eval(mc.density, t);
  return mask_cache_alpha(mc, d) >= mc.thres;
}

// Per-cell verdicts for the mask-cache test.  The interpolated density of a sample in cell (i,j,k) is a convex
// combination of the cell's 8 corner values (ATen's fp32 weights sum to 1 within a few ulp) and alpha(d) is monotone
// in d, so with generous margins for the fp32 evaluation (1e-5 (|d|max + 1) on the density: > 10x the worst-case
// rounding of 8 weight products and 8 accumulations; 1e-5 + 1e-4 thres on alpha: > 10x the error of expf / log1pf /
// 1 - exp cancellation)
//   alpha(min corner - margin) >= thres + tol  =>  every sample of the cell passes   (code 1)
//   alpha(max corner + margin) <= thres - tol  =>  every sample of the cell fails    (code 0)
// and the cell needs the exact evaluation otherwise (code 2: the shell where the mask changes; also the last index along
// each axis, which has no upper corner).  The march then spends the 8 loads + softplus + 2 exp only on the shell, with
// bit-identical keep flags (tests/test_gpu_ops.py::test_march_cell_verdicts_are_exact).
__global__ void k_mask_cache_cells(VxMaskCache mc, uint8_t* __restrict__ cells) {
  const int64_t n = (int64_t)mc.X * mc.Y * mc.Z;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(v % mc.Z), j = (int)((v / mc.Z) % mc.Y), i = (int)(v / ((int64_t)mc.Z * mc.Y));
    uint8_t code = 2;
    if (i < mc.X - 1 && j < mc.Y - 1 && k < mc.Z - 1) {
      float lo = INFINITY, hi = -INFINITY;
      bool finite = true;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float d = mc.density[v + (c & 1) + ((c >> 1) & 1) * mc.Z + (c >> 2) * (int64_t)mc.Y * mc.Z];
        finite &= (fabsf(d) <= 3.0e38f);
        lo = fminf(lo, d); hi = fmaxf(hi, d);
      }
      if (finite) {
        const float margin = 1e-5f * (fmaxf(fabsf(lo), fabsf(hi)) + 1.f);
        const float tol = 1e-5f + 1e-4f * fabsf(mc.thres);
        const float a_lo = mask_cache_alpha(mc, lo - margin), a_hi = mask_cache_alpha(mc, hi + margin);
        if (a_lo >= mc.thres + tol) code = 1;
        else if (a_hi <= mc.thres - tol) code = 0;
      }
    }
    cells[v] = code;
  }
}

VX_API int vx_mask_cache_cells(const float* mc_density, int mc_X, int mc_Y, int mc_Z, float act_shift,
                               float voxel_size_ratio, float thres, uint8_t* cells, cudaStream_t st) {
  const int64_t n = (int64_t)mc_X * mc_Y * mc_Z;
  if (n <= 0) return 0;
  VX_REQUIRE(n < ((int64_t)1 << 31), "vx_mask_cache_cells", "mask grid too large");
  VxMaskCache mc;
  mc.density = mc_density; mc.cells = nullptr; mc.X = mc_X; mc.Y = mc_Y; mc.Z = mc_Z;
  for (int c = 0; c < 3; ++c) { mc.min[c] = 0.f; mc.max[c] = 1.f; }
  mc.act_shift = act_shift; mc.voxel_size_ratio = voxel_size_ratio; mc.thres = thres;
  k_mask_cache_cells<<<(int)min((int64_t)vx_blocks(n, 256), (int64_t)vx_num_sms() * 16), 256, 0, st>>>(mc, cells);
  return vx_check_launch("vx_mask_cache_cells");
}

// standalone MaskCache.forward (voxurf_fine.py:930-942) on explicit points
__global__ void k_mask_cache_query(VxMaskCache mc, const float* __restrict__ xyz, int64_t n, bool* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = mask_cache_keep(mc, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}

VX_API int vx_mask_cache_query(const float* mc_density, int mc_X, int mc_Y, int mc_Z, const float* mc_min_host,
                               const float* mc_max_host, float act_shift, float voxel_size_ratio, float thres,
                               const float* xyz, int64_t n, bool* out, cudaStream_t st) {
  if (n <= 0) return 0;
  VxMaskCache mc;
  mc.density = mc_density; mc.cells = nullptr; mc.X = mc_X; mc.Y = mc_Y; mc.Z = mc_Z;
  for (int c = 0; c < 3; ++c) { mc.min[c] = mc_min_host[c]; mc.max[c] = mc_max_host[c]; }
  mc.act_shift = act_shift; mc.voxel_size_ratio = voxel_size_ratio; mc.thres = thres;
  const int blocks = (int)min((int64_t)vx_blocks(n, 256), (int64_t)vx_num_sms() * 16);
  k_mask_cache_query<<<blocks, 256, 0, st>>>(mc, xyz, n, out);
  return vx_check_launch("vx_mask_cache_query");
}

// pass 1: one warp per ray.  bits: ceil(n_steps/32) words per ray at word offset word_off[r]
// (word_off = exclusive scan of ceil(n/32), computed here from offsets: we simply index words by
// (offsets[r] >> 5) + r, which is a valid injective upper bound because each ray wastes < 1 word).
__global__ void k_march_flags(const float* __restrict__ rays_start, const float* __restrict__ rays_dir,
                              const float* __restrict__ xyz_min, const float* __restrict__ xyz_max,
                              const int64_t* __restrict__ offsets, int n_rays, float stepdist, VxMaskCache mc,
                              uint32_t* __restrict__ bits_inbbox, uint32_t* __restrict__ bits_keep,
                              int* __restrict__ keep_count) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const float mnx = xyz_min[0], mny = xyz_min[1], mnz = xyz_min[2];
  const float mxx = xyz_max[0], mxy = xyz_max[1], mxz = xyz_max[2];
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * warps_per_block) {
    const int64_t beg = offsets[r];
    const int n = (int)(offsets[r + 1] - beg);
    const int64_t w0 = (beg >> 5) + r;
    const float sx = rays_start[3 * r], sy = rays_start[3 * r + 1], sz = rays_start[3 * r + 2];
    const float dx = rays_dir[3 * r], dy = rays_dir[3 * r + 1], dz = rays_dir[3 * r + 2];
    int cnt = 0;
    for (int s0 = 0; s0 < n; s0 += 32) {
      const int s = s0 + lane;
      bool inb = false, keep = false;
      if (s < n) {
        const float dist = stepdist * s;
        const float px = sx + dx * dist;
        const float py = sy + dy * dist;
        const float pz = sz + dz * dist;
        inb = !((mnx > px) | (mny > py) | (mnz > pz) | (mxx < px) | (mxy < py) | (mxz < pz));
        keep = inb && (mc.density == nullptr || mask_cache_keep(mc, px, py, pz));
      }
      const uint32_t bi = __ballot_sync(0xffffffffu, inb);
      const uint32_t bk = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) {
        bits_inbbox[w0 + (s0 >> 5)] = bi;
        bits_keep[w0 + (s0 >> 5)] = bk;
      }
      cnt += __popc(bk);
    }
    if (lane == 0) keep_count[r] = cnt;
  }
}

// exclusive scan of an int array of n (~8192) entries by one CTA; out[n] = total
__global__ void __launch_bounds__(1024) k_scan_i32(const int* __restrict__ in, int n, int* __restrict__ out) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = (i < n) ? in[i] : 0;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += y;
      }
      warp_sum[lane] = w;
    }
    __syncthreads();
    const int incl = x + (warp > 0 ? warp_sum[warp - 1] : 0) + carry_s;
    if (i < n) out[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry_s;
}

VX_API int vx_scan_i32(const int* in, int n, int* out /* n+1 */, cudaStream_t st) {
  if (n < 0) return 0;
  k_scan_i32<<<1, 1024, 0, st>>>(in, n, out);
  return vx_check_launch("vx_scan_i32");
}

// pass 2: expand bit words into the compact list; also (optionally) the legacy M0-sized
// mask_outbbox = !keep, which is what voxurf_fine.py:636 leaves in ret_dict['mask_outbbox'].
__global__ void k_march_emit(const int64_t* __restrict__ offsets, int n_rays, const uint32_t* __restrict__ bits_keep,
                             const int* __restrict__ keep_off, int capacity, int* __restrict__ ray_id,
                             int* __restrict__ step_id, bool* __restrict__ mask_outbbox) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * warps_per_block) {
    const int64_t beg = offsets[r];
    const int n = (int)(offsets[r + 1] - beg);
    const int64_t w0 = (beg >> 5) + r;
    int out = keep_off[r];
    for (int s0 = 0; s0 < n; s0 += 32) {
      const uint32_t bk = bits_keep[w0 + (s0 >> 5)];
      const int s = s0 + lane;
      const bool keep = (bk >> lane) & 1u;
      if (keep) {
        const int dst = out + __popc(bk & ((1u << lane) - 1u));
        if (dst < capacity) { ray_id[dst] = r; step_id[dst] = s; }
      }
      if (mask_outbbox && s < n) mask_outbbox[beg + s] = !keep;
      out += __popc(bk);
    }
  }
}

VX_API int vx_march_flags_cells(const float* rays_start, const float* rays_dir, const float* xyz_min, const float* xyz_max,
                                const int64_t* offsets, int n_rays, float stepdist, const float* mc_density, int mc_X,
                                int mc_Y, int mc_Z, const float* mc_min_host, const float* mc_max_host, float act_shift,
                                float voxel_size_ratio, float thres, const uint8_t* mc_cells, uint32_t* bits_inbbox,
                                uint32_t* bits_keep, int* keep_count, int* keep_off, cudaStream_t st) {
  if (n_rays <= 0) return 0;
  VxMaskCache mc;
  mc.density = mc_density; mc.cells = mc_density ? mc_cells : nullptr; mc.X = mc_X; mc.Y = mc_Y; mc.Z = mc_Z;
  for (int c = 0; c < 3; ++c) { mc.min[c] = mc_density ? mc_min_host[c] : 0.f; mc.max[c] = mc_density ? mc_max_host[c] : 1.f; }
  mc.act_shift = act_shift; mc.voxel_size_ratio = voxel_size_ratio; mc.thres = thres;
  const int blocks = min(vx_blocks((int64_t)n_rays * 32, 256), vx_num_sms() * 8);
  k_march_flags<<<blocks, 256, 0, st>>>(rays_start, rays_dir, xyz_min, xyz_max, offsets, n_rays, stepdist, mc,
                                        bits_inbbox, bits_keep, keep_count);
  int rc = vx_check_launch("vx_march_flags");
  if (rc) return rc;
  k_scan_i32<<<1, 1024, 0, st>>>(keep_count, n_rays, keep_off);
  return vx_check_launch("vx_march_flags(scan)");
}

VX_API int vx_march_flags(const float* rays_start, const float* rays_dir, const float* xyz_min, const float* xyz_max,
                          const int64_t* offsets, int n_rays, float stepdist, const float* mc_density, int mc_X, int mc_Y,
                          int mc_Z, const float* mc_min_host, const float* mc_max_host, float act_shift,
                          float voxel_size_ratio, float thres, uint32_t* bits_inbbox, uint32_t* bits_keep,
                          int* keep_count, int* keep_off, cudaStream_t st) {
  return vx_march_flags_cells(rays_start, rays_dir, xyz_min, xyz_max, offsets, n_rays, stepdist, mc_density, mc_X, mc_Y, mc_Z,
                              mc_min_host, mc_max_host, act_shift, voxel_size_ratio, thres, nullptr, bits_inbbox, bits_keep,
                              keep_count, keep_off, st);
}

VX_API int vx_march_emit(const int64_t* offsets, int n_rays, const uint32_t* bits_keep, const int* keep_off,
                         int capacity, int* ray_id, int* step_id, bool* mask_outbbox, cudaStream_t st) {
  if (n_rays <= 0) return 0;
  const int blocks = min(vx_blocks((int64_t)n_rays * 32, 256), vx_num_sms() * 8);
  k_march_emit<<<blocks, 256, 0, st>>>(offsets, n_rays, bits_keep, keep_off, capacity, ray_id, step_id, mask_outbbox);
  return vx_check_launch("vx_march_emit");
}

// explicit points for a compact (ray_id, step_id) list (the legacy ret-dict wants ray_pts)
__global__ void k_points_from_steps(VxPts src, const int* __restrict__ n_dev, int64_t n_host, float* __restrict__ out) {
  const int64_t n = vx_count(n_dev, n_host);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float px, py, pz;
    vx_load_pt(src, i, px, py, pz);
    out[3 * i] = px; out[3 * i + 1] = py; out[3 * i + 2] = pz;
  }
}

VX_API int vx_points_from_steps(const int* ray_id, const int* step_id, const float* rays_start, const float* rays_dir,
                                float stepdist, const int* n_dev, int64_t n_host, float* out, cudaStream_t st) {
  if (!n_dev && n_host <= 0) return 0;
  VxPts src{nullptr, ray_id, step_id, rays_start, rays_dir, stepdist};
  const int blocks = n_dev ? vx_num_sms() * 8 : min(vx_blocks(n_host, 256), vx_num_sms() * 8);
  k_points_from_steps<<<blocks, 256, 0, st>>>(src, n_dev, n_host, out);
  return vx_check_launch("vx_points_from_steps");
}

// ---------------------------------------------------------------------------------------------
// Caller side of the path: per-view ray generation and the in-mask-cache ray filter.
//   get_rays / ndc_rays / get_rays_of_a_view                      lib/voxurf_fine.py:1001-1067
//   hit_coarse_geo                                                lib/voxurf_fine.py:579-591
//   get_training_rays_in_maskcache_sampling                       lib/voxurf_fine.py:1127-1164
// The reference builds every view with a dozen elementwise ATen ops, tests 64 image rows at a time by materialising
// every sample of every ray (sample_pts_on_rays + MaskCache on ~600 points per ray), and compacts with boolean
// indexing (one host sync per view).  Here: one kernel writes rays_o / rays_d / viewdirs of a view, one warp-per-ray
// kernel marches until the first sample inside the mask (nothing per-sample is stored), one kernel compacts the four
// per-ray arrays of the view behind a running device-side row counter.  One rounding per reference op, same order.
// ---------------------------------------------------------------------------------------------
struct VxCamera {
  float fx, fy, cx, cy;
  float R[3][3];   // c2w[:3,:3]
  float t[3];      // c2w[:3,3]
};

__global__ void k_rays_of_view(int H, int W, VxCamera cam, int inverse_y, int flip_x, int flip_y, float pixel_offset,
                               const float* __restrict__ jitter_i, const float* __restrict__ jitter_j, int ndc, float ndc_cw,
                               float ndc_ch, float ndc_near, float* __restrict__ rays_o, float* __restrict__ rays_d,
                               float* __restrict__ viewdirs) {
  const int n = H * W;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < n; pix += gridDim.x * blockDim.x) {
    const int r = pix / W, c = pix - r * W;
    const int cc = flip_x ? W - 1 - c : c, rr = flip_y ? H - 1 - r : r;     // i.flip((1,)), j.flip((0,))
    float i = (float)cc, j = (float)rr;
    if (jitter_i) { i = __fadd_rn(i, jitter_i[r * W + cc]); j = __fadd_rn(j, jitter_j[rr * W + c]); }   // mode 'random'
    else { i = __fadd_rn(i, pixel_offset); j = __fadd_rn(j, pixel_offset); }                            // 'center' / 'lefttop'
    float d0 = __fdiv_rn(__fsub_rn(i, cam.cx), cam.fx);
    float d1 = __fdiv_rn(__fsub_rn(j, cam.cy), cam.fy);
    float d2 = 1.f;
    if (!inverse_y) { d1 = -d1; d2 = -1.f; }
    float o[3], d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {   // torch.sum(dirs[..., None, :] * c2w[:3,:3], -1)
      d[k] = __fadd_rn(__fadd_rn(__fmul_rn(d0, cam.R[k][0]), __fmul_rn(d1, cam.R[k][1])), __fmul_rn(d2, cam.R[k][2]));
      o[k] = cam.t[k];
    }
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
#pragma unroll
    for (int k = 0; k < 3; ++k) viewdirs[3 * pix + k] = __fdiv_rn(d[k], nrm);
    if (ndc) {   // ndc_rays(H, W, focal, near = 1., rays_o, rays_d), lib/voxurf_fine.py:1037-1056
      const float t = __fdiv_rn(-__fadd_rn(ndc_near, o[2]), d[2]);
#pragma unroll
      for (int k = 0; k < 3; ++k) o[k] = __fadd_rn(o[k], __fmul_rn(t, d[k]));
      const float o0 = __fdiv_rn(__fmul_rn(ndc_cw, o[0]), o[2]);
      const float o1 = __fdiv_rn(__fmul_rn(ndc_ch, o[1]), o[2]);
      const float o2 = __fadd_rn(1.f, __fdiv_rn(__fmul_rn(2.f, ndc_near), o[2]));
      const float e0 = __fmul_rn(ndc_cw, __fsub_rn(__fdiv_rn(d[0], d[2]), __fdiv_rn(o[0], o[2])));
      const float e1 = __fmul_rn(ndc_ch, __fsub_rn(__fdiv_rn(d[1], d[2]), __fdiv_rn(o[1], o[2])));
      const float e2 = __fdiv_rn(__fmul_rn(-2.f, ndc_near), o[2]);
      o[0] = o0; o[1] = o1; o[2] = o2; d[0] = e0; d[1] = e1; d[2] = e2;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { rays_o[3 * pix + k] = o[k]; rays_d[3 * pix + k] = d[k]; }
  }
}

// K_host: 3x3 row-major intrinsics; c2w_host: the first three rows of the pose, 3x4 row-major.
// mode: 0 'lefttop', 1 'center', 2 'random' (needs jitter_i, jitter_j: (H,W) uniform [0,1) offsets).
VX_API int vx_rays_of_view(int H, int W, const float* K_host, const float* c2w_host, int inverse_y, int flip_x, int flip_y,
                           int mode, const float* jitter_i, const float* jitter_j, int ndc, float ndc_near, float* rays_o,
                           float* rays_d, float* viewdirs, cudaStream_t st) {
  VX_REQUIRE(H >= 0 && W >= 0 && (int64_t)H * W < ((int64_t)1 << 30), "vx_rays_of_view", "image size");
  VX_REQUIRE(mode >= 0 && mode <= 2, "vx_rays_of_view", "mode must be 0 (lefttop), 1 (center) or 2 (random)");
  VX_REQUIRE(mode != 2 || (jitter_i && jitter_j), "vx_rays_of_view", "mode 'random' needs the jitter arrays");
  if (H * W == 0) return 0;
  VxCamera cam;
  cam.fx = K_host[0]; cam.cx = K_host[2]; cam.fy = K_host[4]; cam.cy = K_host[5];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) cam.R[r][c] = c2w_host[4 * r + c];
    cam.t[r] = c2w_host[4 * r + 3];
  }
  // -1./(W/(2.*focal)) evaluated in double like the Python expression, then narrowed once
  const double focal = (double)K_host[0];
  const float cw = (float)(-1.0 / ((double)W / (2.0 * focal))), ch = (float)(-1.0 / ((double)H / (2.0 * focal)));
  const int blocks = min(vx_blocks((int64_t)H * W, 256), vx_num_sms() * 16);
  k_rays_of_view<<<blocks, 256, 0, st>>>(H, W, cam, inverse_y, flip_x, flip_y, mode == 1 ? 0.5f : 0.f,
                                         mode == 2 ? jitter_i : nullptr, mode == 2 ? jitter_j : nullptr, ndc, cw, ch,
                                         ndc_near, rays_o, rays_d, viewdirs);
  return vx_check_launch("vx_rays_of_view");
}

// hit[r] = any sample of ray r lies inside the bbox and inside the mask cache; one warp per ray, 32 steps per round,
// stops at the first round with a hit
__global__ void k_rays_hit(const float* __restrict__ rays_o, const float* __restrict__ rays_d, int n_rays, VxGrid box,
                           float near, float far, float stepdist, VxMaskCache mc, bool* __restrict__ hit) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * warps_per_block) {
    const float* o = rays_o + 3 * r;
    const float* d = rays_d + 3 * r;
    float tmn, tmx;
    ray_t_minmax(o, d, box.min, box.max, near, far, tmn, tmx);
    const float rnorm = ray_norm(d);
    const int n = (int)ray_n_samples(tmn, tmx, rnorm, stepdist);
    const float sx = o[0] + d[0] * tmn, sy = o[1] + d[1] * tmn, sz = o[2] + d[2] * tmn;
    const float dx = d[0] / rnorm, dy = d[1] / rnorm, dz = d[2] / rnorm;
    bool any = false;
    for (int s0 = 0; s0 < n && !any; s0 += 32) {
      const int s = s0 + lane;
      bool keep = false;
      if (s < n) {
        const float dist = stepdist * s;
        const float px = sx + dx * dist;
        const float py = sy + dy * dist;
        const float pz = sz + dz * dist;
        const bool inb = !((box.min[0] > px) | (box.min[1] > py) | (box.min[2] > pz) | (box.max[0] < px) | (box.max[1] < py) |
                           (box.max[2] < pz));
        keep = inb && mask_cache_keep(mc, px, py, pz);
      }
      any = __any_sync(0xffffffffu, keep);
    }
    if (lane == 0) hit[r] = any;
  }
}

VX_API int vx_rays_hit_mask(const float* rays_o, const float* rays_d, int n_rays, const float* xyz_min_host,
                            const float* xyz_max_host, float near, float far, float stepdist, const float* mc_density,
                            int mc_X, int mc_Y, int mc_Z, const float* mc_min_host, const float* mc_max_host,
                            float act_shift, float voxel_size_ratio, float thres, bool* hit, cudaStream_t st) {
  VX_REQUIRE(mc_density != nullptr, "vx_rays_hit_mask", "needs a mask cache");
  if (n_rays <= 0) return 0;
  VxGrid box;
  box.X = box.Y = box.Z = box.C = 1; box.cl = 0;
  VxMaskCache mc;
  mc.density = mc_density; mc.cells = nullptr; mc.X = mc_X; mc.Y = mc_Y; mc.Z = mc_Z;
  for (int c = 0; c < 3; ++c) {
    box.min[c] = xyz_min_host[c]; box.max[c] = xyz_max_host[c];
    mc.min[c] = mc_min_host[c]; mc.max[c] = mc_max_host[c];
  }
  mc.act_shift = act_shift; mc.voxel_size_ratio = voxel_size_ratio; mc.thres = thres;
  const int blocks = min(vx_blocks((int64_t)n_rays * 32, 256), vx_num_sms() * 16);
  k_rays_hit<<<blocks, 256, 0, st>>>(rays_o, rays_d, n_rays, box, near, far, stepdist, mc, hit);
  return vx_check_launch("vx_rays_hit_mask");
}

// stable compaction of the rows of up to four (n,3) arrays selected by `mask`, appended at row tops[k]; writes
// tops[k+1] = tops[k] + count.  incl = inclusive prefix sum of mask (int32).
__global__ void k_compact_rows3(const bool* __restrict__ mask, const int* __restrict__ incl, int n,
                                const int64_t* __restrict__ top_in, int64_t* __restrict__ top_out, int64_t capacity,
                                const float* __restrict__ s0, const float* __restrict__ s1, const float* __restrict__ s2,
                                const float* __restrict__ s3, float* __restrict__ d0, float* __restrict__ d1,
                                float* __restrict__ d2, float* __restrict__ d3) {
  const int64_t top = *top_in;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (i == n - 1) *top_out = top + incl[i];
    if (!mask[i]) continue;
    const int64_t dst = top + incl[i] - 1;
    if (dst >= capacity) continue;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (s0) d0[3 * dst + k] = s0[3 * (int64_t)i + k];
      if (s1) d1[3 * dst + k] = s1[3 * (int64_t)i + k];
      if (s2) d2[3 * dst + k] = s2[3 * (int64_t)i + k];
      if (s3) d3[3 * dst + k] = s3[3 * (int64_t)i + k];
    }
  }
}

VX_API int vx_compact_rows3(const bool* mask, const int* incl, int n, const int64_t* top_in, int64_t* top_out,
                            int64_t capacity, const float* src0, const float* src1, const float* src2, const float* src3,
                            float* dst0, float* dst1, float* dst2, float* dst3, cudaStream_t st) {
  VX_REQUIRE(top_in && top_out && top_in != top_out, "vx_compact_rows3", "top_in / top_out must be distinct device slots");
  if (n <= 0) return 0;
  k_compact_rows3<<<min(vx_blocks(n, 256), vx_num_sms() * 16), 256, 0, st>>>(mask, incl, n, top_in, top_out, capacity, src0,
                                                                           src1, src2, src3, dst0, dst1, dst2, dst3);
  return vx_check_launch("vx_compact_rows3");
}
