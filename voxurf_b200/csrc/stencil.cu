// Grid-level stencils (SURVEY.md 8a rows A13, A14, A20 and the smooth-gradient TV regulariser).
//   * finite-difference SDF gradient grid, fwd + adjoint      lib/voxurf_fine.py:440-460
//   * Gaussian / binomial smoothing Conv3d(k, replicate pad)   lib/voxurf_fine.py:246-258, 236-242
//     fwd + adjoint (the coarse stage back-propagates through it every iteration, voxurf_coarse.py:531)
//   * smooth-gradient TV loss                                 lib/voxurf_fine.py:412-421
//   * total_variation_add_grad / _new                         lib/cuda/total_variation_kernel.cu:14-133
// All are HBM-bound streaming kernels over (X,Y,Z) grids with Z fastest: threads map to
// consecutive Z so every access of a warp is one or two 128-byte lines; neighbour planes come
// from L1/L2.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// central-difference gradient grid ('interpolate' mode): zero on the two boundary planes of an axis
// ---------------------------------------------------------------------------------------------
__global__ void k_fd_gradient(const float* __restrict__ sdf, int X, int Y, int Z, float voxel_size,
                              float* __restrict__ grad /* (3,X,Y,Z) */) {
  const int64_t V = (int64_t)X * Y * Z;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
    const int k = v % Z, j = (v / Z) % Y, i = v / ((int64_t)Z * Y);
    const int64_t sX = (int64_t)Y * Z;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (i > 0 && i < X - 1) gx = __fdiv_rn(__fdiv_rn(__fsub_rn(__ldg(sdf + v + sX), __ldg(sdf + v - sX)), 2.f), voxel_size);
    if (j > 0 && j < Y - 1) gy = __fdiv_rn(__fdiv_rn(__fsub_rn(__ldg(sdf + v + Z), __ldg(sdf + v - Z)), 2.f), voxel_size);
    if (k > 0 && k < Z - 1) gz = __fdiv_rn(__fdiv_rn(__fsub_rn(__ldg(sdf + v + 1), __ldg(sdf + v - 1)), 2.f), voxel_size);
    grad[v] = gx; grad[V + v] = gy; grad[2 * V + v] = gz;
  }
}

// adjoint: dsdf[u] += dG_x[u - eX]/(2 vs) (if u-eX interior in x) - dG_x[u + eX]/(2 vs) (if interior) + ...
__global__ void k_fd_gradient_bwd(const float* __restrict__ dgrad, int X, int Y, int Z, float voxel_size,
                                  float* __restrict__ dsdf) {
  const int64_t V = (int64_t)X * Y * Z;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
    const int k = v % Z, j = (v / Z) % Y, i = v / ((int64_t)Z * Y);
    const int64_t sX = (int64_t)Y * Z;
    float acc = 0.f;
    if (i - 1 > 0 && i - 1 < X - 1) acc += __ldg(dgrad + v - sX);
    if (i + 1 > 0 && i + 1 < X - 1) acc -= __ldg(dgrad + v + sX);
    if (j - 1 > 0 && j - 1 < Y - 1) acc += __ldg(dgrad + V + v - Z);
    if (j + 1 > 0 && j + 1 < Y - 1) acc -= __ldg(dgrad + V + v + Z);
    if (k - 1 > 0 && k - 1 < Z - 1) acc += __ldg(dgrad + 2 * V + v - 1);
    if (k + 1 > 0 && k + 1 < Z - 1) acc -= __ldg(dgrad + 2 * V + v + 1);
    dsdf[v] += (acc / voxel_size) * 0.5f;
  }
}

// ---------------------------------------------------------------------------------------------
// Vectorised forms (Z % 4 == 0, numel < 2^31, 16-byte aligned): one thread = 4 consecutive z, 32-bit index
// arithmetic, 128-bit loads / stores.  Every output element is produced by exactly the scalar kernels' sequence of
// rounded operations, so the two forms agree bit for bit (x / 2 == x * 0.5 exactly).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float fd1(float hi, float lo, float vs) { return __fdiv_rn(__fmul_rn(__fsub_rn(hi, lo), 0.5f), vs); }
__device__ __forceinline__ float4 fd4(const float4& a, const float4& b, float vs) {
  return make_float4(fd1(a.x, b.x, vs), fd1(a.y, b.y, vs), fd1(a.z, b.z, vs), fd1(a.w, b.w, vs));
}
static bool vec4_ok(int X, int Y, int Z, int64_t planes, const void* a, const void* b, const void* c = nullptr) {
  const int64_t n = (int64_t)X * Y * Z * planes;
  return Z % 4 == 0 && n < ((int64_t)1 << 31) &&
         ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}

// `active` (optional, one byte per voxel): groups of 4 voxels with no active voxel are skipped -- their entries of grad
// are left as they are.  The smooth-gradient TV only reads the gradient within one voxel of the non-empty mask.
__global__ void __launch_bounds__(256) k_fd_gradient_v4(const float* __restrict__ sdf, uint32_t X, uint32_t Y, uint32_t Z,
                                                        float vs, float* __restrict__ grad, const bool* __restrict__ active) {
  const uint32_t Z4 = Z >> 2, n4 = X * Y * Z4, V = X * Y * Z, sX = Y * Z;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += gridDim.x * blockDim.x) {
    if (active && __ldg(reinterpret_cast<const uint32_t*>(active) + t) == 0u) continue;
    const uint32_t k4 = t % Z4, ij = t / Z4, j = ij % Y, i = ij / Y, v = t << 2;
    const float4 c = ld4(sdf + v);
    float4 gx = make_float4(0.f, 0.f, 0.f, 0.f), gy = gx, gz;
    if (i > 0 && i < X - 1) gx = fd4(ld4(sdf + v + sX), ld4(sdf + v - sX), vs);
    if (j > 0 && j < Y - 1) gy = fd4(ld4(sdf + v + Z), ld4(sdf + v - Z), vs);
    gz.x = k4 > 0 ? fd1(c.y, __ldg(sdf + v - 1), vs) : 0.f;
    gz.y = fd1(c.z, c.x, vs);
    gz.z = fd1(c.w, c.y, vs);
    gz.w = k4 < Z4 - 1 ? fd1(__ldg(sdf + v + 4), c.z, vs) : 0.f;
    *reinterpret_cast<float4*>(grad + v) = gx;
    *reinterpret_cast<float4*>(grad + V + v) = gy;
    *reinterpret_cast<float4*>(grad + 2 * V + v) = gz;
  }
}

// grad += FD^T(dgrad) (vx_fd_gradient_backward), then, when kTv, grad += the dense unmasked total_variation_add_grad
// of `param` (total_variation_kernel.cu:14-35) -- one read-modify-write of grad for both regularisers.
// `active` (optional): where no voxel of the group of 4 is active, dgrad is known to be zero at all six neighbours and the
// FD part is skipped.
template <bool kFd, bool kTv>
__global__ void __launch_bounds__(256) k_sdf_reg_backward_v4(const float* __restrict__ dgrad, const float* __restrict__ param,
                                                             uint32_t X, uint32_t Y, uint32_t Z, float vs, float w_fast,
                                                             float w_mid, float w_slow, float* __restrict__ grad,
                                                             const bool* __restrict__ active, uint32_t x_begin = 0,
                                                             uint32_t x_end = 0xffffffffu) {
  // [x_begin, x_end): the X-slab of grad to update (data-parallel training: every rank regularises and steps only the
  // slab it owns; dgrad / param are the full grids, the one-plane halo is local)
  const uint32_t Z4 = Z >> 2, V = X * Y * Z, sX = Y * Z;
  const uint32_t t0 = x_begin * Y * Z4, n4 = min(x_end, X) * Y * Z4;
  for (uint32_t t = t0 + blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += gridDim.x * blockDim.x) {
    const uint32_t k4 = t % Z4, ij = t / Z4, j = ij % Y, i = ij / Y, v = t << 2, k = k4 << 2;
    const bool fd_here = kFd && !(active && __ldg(reinterpret_cast<const uint32_t*>(active) + t) == 0u);
    if (!kTv && !fd_here) continue;
    float4 g = *reinterpret_cast<const float4*>(grad + v);
    float* gp = reinterpret_cast<float*>(&g);
    if (fd_here) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      // same order as k_fd_gradient_bwd: x-, x+, y-, y+, z-, z+ ; a source contributes iff it is interior on its axis
      if (i >= 2 && i - 1 < X - 1) { const float4 a = ld4(dgrad + v - sX); acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w; }
      if (i + 1 < X - 1) { const float4 a = ld4(dgrad + v + sX); acc[0] -= a.x; acc[1] -= a.y; acc[2] -= a.z; acc[3] -= a.w; }
      if (j >= 2 && j - 1 < Y - 1) { const float4 a = ld4(dgrad + V + v - Z); acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w; }
      if (j + 1 < Y - 1) { const float4 a = ld4(dgrad + V + v + Z); acc[0] -= a.x; acc[1] -= a.y; acc[2] -= a.z; acc[3] -= a.w; }
      const float4 cz = ld4(dgrad + 2 * V + v);
      const float zl = k4 > 0 ? __ldg(dgrad + 2 * V + v - 1) : 0.f, zr = k4 < Z4 - 1 ? __ldg(dgrad + 2 * V + v + 4) : 0.f;
      const float zv[6] = {zl, cz.x, cz.y, cz.z, cz.w, zr};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t kk = k + q;
        if (kk >= 2 && kk - 1 < Z - 1) acc[q] += zv[q];
        if (kk + 1 < Z - 1) acc[q] -= zv[q + 2];
        gp[q] += (acc[q] / vs) * 0.5f;
      }
    }
    if (kTv) {
      const float4 c = ld4(param + v);
      const float pl = k4 > 0 ? __ldg(param + v - 1) : 0.f, pr = k4 < Z4 - 1 ? __ldg(param + v + 4) : 0.f;
      const float pv[6] = {pl, c.x, c.y, c.z, c.w, pr};
      float ym[4] = {0.f, 0.f, 0.f, 0.f}, yp[4] = {0.f, 0.f, 0.f, 0.f}, xm[4] = {0.f, 0.f, 0.f, 0.f}, xp[4] = {0.f, 0.f, 0.f, 0.f};
      if (j > 0) { const float4 a = ld4(param + v - Z); ym[0] = a.x; ym[1] = a.y; ym[2] = a.z; ym[3] = a.w; }
      if (j < Y - 1) { const float4 a = ld4(param + v + Z); yp[0] = a.x; yp[1] = a.y; yp[2] = a.z; yp[3] = a.w; }
      if (i > 0) { const float4 a = ld4(param + v - sX); xm[0] = a.x; xm[1] = a.y; xm[2] = a.z; xm[3] = a.w; }
      if (i < X - 1) { const float4 a = ld4(param + v + sX); xp[0] = a.x; xp[1] = a.y; xp[2] = a.z; xp[3] = a.w; }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t kk = k + q;
        const float p0 = pv[q + 1];
        float add = 0;
        if (kk != 0) add += w_fast * fminf(fmaxf(p0 - pv[q], -1.f), 1.f);
        if (kk != Z - 1) add += w_fast * fminf(fmaxf(p0 - pv[q + 2], -1.f), 1.f);
        if (j != 0) add += w_mid * fminf(fmaxf(p0 - ym[q], -1.f), 1.f);
        if (j != Y - 1) add += w_mid * fminf(fmaxf(p0 - yp[q], -1.f), 1.f);
        if (i != 0) add += w_slow * fminf(fmaxf(p0 - xm[q], -1.f), 1.f);
        if (i != X - 1) add += w_slow * fminf(fmaxf(p0 - xp[q], -1.f), 1.f);
        gp[q] += add;
      }
    }
    *reinterpret_cast<float4*>(grad + v) = g;
  }
}

static int grid_blocks(int64_t V) { return (int)min((int64_t)vx_blocks(V, 256), (int64_t)vx_num_sms() * 16); }

VX_API int vx_fd_gradient(const float* sdf, int X, int Y, int Z, float voxel_size, float* grad, cudaStream_t st) {
  const int64_t V = (int64_t)X * Y * Z;
  if (V <= 0) return 0;
  if (vec4_ok(X, Y, Z, 3, sdf, grad) && V % 4 == 0) {   // (the y/z component planes start at multiples of V floats)
    k_fd_gradient_v4<<<grid_blocks(V / 4), 256, 0, st>>>(sdf, X, Y, Z, voxel_size, grad, nullptr);
    return vx_check_launch("vx_fd_gradient");
  }
  k_fd_gradient<<<grid_blocks(V), 256, 0, st>>>(sdf, X, Y, Z, voxel_size, grad);
  return vx_check_launch("vx_fd_gradient");
}

// vx_fd_gradient restricted to the groups of 4 z-consecutive voxels that contain an active voxel; the rest of grad is
// left untouched (callers keep it zero-initialised).  Falls back to the full pass when the vectorised form does not apply.
VX_API int vx_fd_gradient_active(const float* sdf, int X, int Y, int Z, float voxel_size, const bool* active, float* grad,
                                 cudaStream_t st) {
  const int64_t V = (int64_t)X * Y * Z;
  if (V <= 0) return 0;
  if (active && vec4_ok(X, Y, Z, 3, sdf, grad) && V % 4 == 0 && (reinterpret_cast<uintptr_t>(active) & 3) == 0) {
    k_fd_gradient_v4<<<grid_blocks(V / 4), 256, 0, st>>>(sdf, X, Y, Z, voxel_size, grad, active);
    return vx_check_launch("vx_fd_gradient_active");
  }
  return vx_fd_gradient(sdf, X, Y, Z, voxel_size, grad, st);
}

VX_API int vx_fd_gradient_backward(const float* dgrad, int X, int Y, int Z, float voxel_size, float* dsdf,
                                   cudaStream_t st) {
  const int64_t V = (int64_t)X * Y * Z;
  if (V <= 0) return 0;
  if (vec4_ok(X, Y, Z, 3, dgrad, dsdf)) {
    k_sdf_reg_backward_v4<true, false><<<grid_blocks(V / 4), 256, 0, st>>>(dgrad, nullptr, X, Y, Z, voxel_size, 0.f, 0.f, 0.f, dsdf, nullptr);
    return vx_check_launch("vx_fd_gradient_backward");
  }
  k_fd_gradient_bwd<<<grid_blocks(V), 256, 0, st>>>(dgrad, X, Y, Z, voxel_size, dsdf);
  return vx_check_launch("vx_fd_gradient_backward");
}

// ---------------------------------------------------------------------------------------------
// k^3 convolution with replicate padding over B independent (X,Y,Z) volumes, k in {3,5};
// weights (k,k,k) in constant-cache friendly kernel-argument space.
// ---------------------------------------------------------------------------------------------
struct VxKernel3 {
  int k;
  float w[125];
};

__global__ void k_conv3d_replicate(const float* __restrict__ in, int B, int X, int Y, int Z, VxKernel3 ker,
                                   float* __restrict__ out) {
  const int64_t V = (int64_t)X * Y * Z;
  const int p = ker.k / 2;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < V * B; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = t % V;
    const float* src = in + (t / V) * V;
    const int z = v % Z, y = (v / Z) % Y, x = v / ((int64_t)Z * Y);
    float acc = 0.f;
    for (int a = 0; a < ker.k; ++a) {
      const int xx = min(max(x + a - p, 0), X - 1);
      for (int b = 0; b < ker.k; ++b) {
        const int yy = min(max(y + b - p, 0), Y - 1);
        const float* row = src + ((int64_t)xx * Y + yy) * Z;
        for (int c = 0; c < ker.k; ++c) {
          const int zz = min(max(z + c - p, 0), Z - 1);
          acc += __ldg(row + zz) * ker.w[(a * ker.k + b) * ker.k + c];
        }
      }
    }
    out[t] = acc;
  }
}

// adjoint (gather form): din[u] = sum_o w[o] * sum_{v : clamp(v + o) == u} dout[v].  Per axis the
// pre-image of u under v -> clamp(v + o) is the single point u - o for interior u and a short
// range on the two boundary planes.
__device__ __forceinline__ void preimage(int u, int o, int n, int& lo, int& hi) {
  if (u > 0 && u < n - 1) { lo = hi = u - o; if (lo < 0 || lo > n - 1) { lo = 1; hi = 0; } }
  else if (n == 1) { lo = 0; hi = 0; }
  else if (u == 0) { lo = 0; hi = min(-o, n - 1); }           // v + o <= 0
  else { lo = max(n - 1 - o, 0); hi = n - 1; }                // v + o >= n-1
}

__global__ void k_conv3d_replicate_bwd(const float* __restrict__ dout, int B, int X, int Y, int Z, VxKernel3 ker,
                                       int accumulate, float* __restrict__ din) {
  const int64_t V = (int64_t)X * Y * Z;
  const int p = ker.k / 2;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < V * B; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = t % V;
    const float* src = dout + (t / V) * V;
    const int z = v % Z, y = (v / Z) % Y, x = v / ((int64_t)Z * Y);
    float acc = 0.f;
    for (int a = 0; a < ker.k; ++a) {
      int xl, xh;
      preimage(x, a - p, X, xl, xh);
      for (int b = 0; b < ker.k; ++b) {
        int yl, yh;
        preimage(y, b - p, Y, yl, yh);
        for (int c = 0; c < ker.k; ++c) {
          int zl, zh;
          preimage(z, c - p, Z, zl, zh);
          float s = 0.f;
          for (int xx = xl; xx <= xh; ++xx)
            for (int yy = yl; yy <= yh; ++yy)
              for (int zz = zl; zz <= zh; ++zz) s += __ldg(src + ((int64_t)xx * Y + yy) * Z + zz);
          acc += s * ker.w[(a * ker.k + b) * ker.k + c];
        }
      }
    }
    din[t] = accumulate ? din[t] + acc : acc;
  }
}

static int fill_kernel(VxKernel3& k, const float* w_host, int ksize) {
  if (ksize != 1 && ksize != 3 && ksize != 5) return -1;
  k.k = ksize;
  for (int i = 0; i < ksize * ksize * ksize; ++i) k.w[i] = w_host[i];
  return 0;
}

VX_API int vx_conv3d_replicate(const float* in, int B, int X, int Y, int Z, const float* weight_host, int ksize,
                               float* out, cudaStream_t st) {
  VxKernel3 ker;
  VX_REQUIRE(fill_kernel(ker, weight_host, ksize) == 0, "vx_conv3d_replicate", "ksize must be 1, 3 or 5");
  const int64_t n = (int64_t)X * Y * Z * B;
  if (n <= 0) return 0;
  k_conv3d_replicate<<<grid_blocks(n), 256, 0, st>>>(in, B, X, Y, Z, ker, out);
  return vx_check_launch("vx_conv3d_replicate");
}

VX_API int vx_conv3d_replicate_backward(const float* dout, int B, int X, int Y, int Z, const float* weight_host,
                                        int ksize, int accumulate, float* din, cudaStream_t st) {
  VxKernel3 ker;
  VX_REQUIRE(fill_kernel(ker, weight_host, ksize) == 0, "vx_conv3d_replicate_backward", "ksize must be 1, 3 or 5");
  const int64_t n = (int64_t)X * Y * Z * B;
  if (n <= 0) return 0;
  k_conv3d_replicate_bwd<<<grid_blocks(n), 256, 0, st>>>(dout, B, X, Y, Z, ker, accumulate, din);
  return vx_check_launch("vx_conv3d_replicate_backward");
}

// ---------------------------------------------------------------------------------------------
// Separable form of the same convolution.  The reference's smoothing kernel is a normalised Gaussian
// exp(-(x^2+y^2+z^2) / 2 sigma^2) / sum (lib/voxurf_fine.py:246-254) = g (x) g (x) g with g the normalised 1-D
// Gaussian, and replicate padding commutes with the factorisation, so three 1-D passes of k taps (15 taps for k = 5)
// replace k^3 = 125; the adjoint is the three 1-D adjoints.  Z pass in -> out, Y pass out -> scratch, X pass
// scratch -> out.  Agreement with the k^3 form is at rounding level (different summation order).
// ---------------------------------------------------------------------------------------------
struct VxKernel1 {
  int k;
  float w[5];
};

template <bool kAdjoint>
__global__ void __launch_bounds__(256) k_conv1d_replicate(const float* __restrict__ in, float* __restrict__ out, uint32_t n_total,
                                                          uint32_t n_axis, uint32_t stride, VxKernel1 ker, int accumulate) {
  const int p = ker.k / 2, n = (int)n_axis;
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n_total; v += gridDim.x * blockDim.x) {
    const int a = (int)((v / stride) % n_axis);
    const float* line = in + v - (uint32_t)a * stride;        // element 0 of this voxel's line along the axis
    float acc = 0.f;
    if (!kAdjoint) {
      for (int o = 0; o < ker.k; ++o) acc += __ldg(line + (uint32_t)min(max(a + o - p, 0), n - 1) * stride) * ker.w[o];
    } else {
      // din[a] = sum_d w[d + p] * sum_{v : clamp(v + d) == a} dout[v]
      for (int o = 0; o < ker.k; ++o) {
        const int d = o - p;
        int lo, hi;
        if (n == 1) { lo = 0; hi = 0; }
        else if (a > 0 && a < n - 1) { lo = hi = a - d; if (lo < 0 || lo > n - 1) { lo = 1; hi = 0; } }
        else if (a == 0) { lo = 0; hi = min(-d, n - 1); }
        else { lo = max(n - 1 - d, 0); hi = n - 1; }
        float sm = 0.f;
        for (int u = lo; u <= hi; ++u) sm += __ldg(line + (uint32_t)u * stride);
        acc += sm * ker.w[o];
      }
    }
    out[v] = accumulate ? out[v] + acc : acc;
  }
}

// w1_host: the k 1-D weights.  adjoint = 0: out (+)= conv(in); adjoint = 1: out (+)= conv^T(in).
// scratch: 2 * B*X*Y*Z floats (the two intermediate volumes); `in` is not modified, `out` may alias nothing else.
VX_API int vx_conv3d_replicate_separable(const float* in, int B, int X, int Y, int Z, const float* w1_host, int ksize, int adjoint,
                                         int accumulate, float* scratch, float* out, cudaStream_t st) {
  VX_REQUIRE(ksize == 1 || ksize == 3 || ksize == 5, "vx_conv3d_replicate_separable", "ksize must be 1, 3 or 5");
  const int64_t n64 = (int64_t)B * X * Y * Z;
  if (n64 <= 0) return 0;
  VX_REQUIRE(n64 < ((int64_t)1 << 31), "vx_conv3d_replicate_separable", "more than 2^31 elements");
  VX_REQUIRE(scratch && scratch != in && scratch != out && in != out, "vx_conv3d_replicate_separable", "distinct in / scratch / out");
  VxKernel1 ker;
  ker.k = ksize;
  for (int i = 0; i < 5; ++i) ker.w[i] = i < ksize ? w1_host[i] : 0.f;
  const uint32_t n = (uint32_t)n64;
  float* s0 = scratch;
  float* s1 = scratch + n64;
  const int blocks = grid_blocks(n64);
  const uint32_t sZ = 1u, sY = (uint32_t)Z, sX = (uint32_t)Y * (uint32_t)Z;
  if (!adjoint) {
    k_conv1d_replicate<false><<<blocks, 256, 0, st>>>(in, s0, n, (uint32_t)Z, sZ, ker, 0);
    k_conv1d_replicate<false><<<blocks, 256, 0, st>>>(s0, s1, n, (uint32_t)Y, sY, ker, 0);
    k_conv1d_replicate<false><<<blocks, 256, 0, st>>>(s1, out, n, (uint32_t)X, sX, ker, accumulate);
  } else {
    k_conv1d_replicate<true><<<blocks, 256, 0, st>>>(in, s0, n, (uint32_t)X, sX, ker, 0);
    k_conv1d_replicate<true><<<blocks, 256, 0, st>>>(s0, s1, n, (uint32_t)Y, sY, ker, 0);
    k_conv1d_replicate<true><<<blocks, 256, 0, st>>>(s1, out, n, (uint32_t)Z, sZ, ker, accumulate);
  }
  return vx_check_launch("vx_conv3d_replicate_separable");
}

// ---------------------------------------------------------------------------------------------
// smooth-gradient TV (lib/voxurf_fine.py:417-420):
//   E_a = tv_smooth_conv(G_a).detach() - G_a ;  loss = w * mean(E[mask x 3]^2)
// pass 1 (this kernel): E from the materialised FD gradient G (3,X,Y,Z); writes dL/dG (3,X,Y,Z)
//   = -2 * w / (3 * n_mask) * E * mask  and per-block partial sums of E^2 (deterministic 2-stage sum).
// pass 2: vx_fd_gradient_backward pushes dL/dG into the sdf gradient.
// ---------------------------------------------------------------------------------------------
__global__ void k_smooth_grad_tv(const float* __restrict__ G, const bool* __restrict__ mask, int X, int Y, int Z,
                                 VxKernel3 ker, float scale /* w / (3 n_mask) */, float* __restrict__ dG,
                                 float* __restrict__ partial) {
  __shared__ float red[32];
  const int64_t V = (int64_t)X * Y * Z;
  float local = 0.f;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < 3 * V; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = t % V;
    float d = 0.f;
    if (mask[v]) {
      const float* src = G + (t / V) * V;
      const int z = v % Z, y = (v / Z) % Y, x = v / ((int64_t)Z * Y);
      float acc = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const int xx = min(max(x + a - 1, 0), X - 1);
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const int yy = min(max(y + b - 1, 0), Y - 1);
          const float* row = src + ((int64_t)xx * Y + yy) * Z;
#pragma unroll
          for (int c = 0; c < 3; ++c) acc += __ldg(row + min(max(z + c - 1, 0), Z - 1)) * ker.w[(a * 3 + b) * 3 + c];
        }
      }
      const float e = acc - __ldg(src + v);
      local += e * e;
      d = -2.f * scale * e;
    }
    dG[t] = d;
  }
  // block reduction in a fixed order
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256) k_smooth_grad_tv_v4(const float* __restrict__ G, const bool* __restrict__ mask,
                                                          uint32_t X, uint32_t Y, uint32_t Z, VxKernel3 ker, float scale,
                                                          float* __restrict__ dG, float* __restrict__ partial, int skip_unmasked) {
  __shared__ float red[32];
  const uint32_t Z4 = Z >> 2, V = X * Y * Z, V4 = V >> 2;
  float local = 0.f;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < 3 * V4; t += gridDim.x * blockDim.x) {
    const uint32_t comp = t / V4, u4 = t - comp * V4, v = u4 << 2;
    const uint32_t k4 = u4 % Z4, ij = u4 / Z4, y = ij % Y, x = ij / Y;
    const uchar4 mk = __ldg(reinterpret_cast<const uchar4*>(mask) + u4);
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!(mk.x | mk.y | mk.z | mk.w)) {
      if (skip_unmasked) continue;     // the caller keeps dG zero there (static mask, zero-initialised buffer)
    } else {
      const float* src = G + comp * V;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      float ctr[4];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const uint32_t xx = min(max((int)x + a - 1, 0), (int)X - 1);
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const uint32_t yy = min(max((int)y + b - 1, 0), (int)Y - 1);
          const float* row = src + (xx * Y + yy) * Z + (k4 << 2);
          const float4 m = ld4(row);
          const float l = k4 > 0 ? __ldg(row - 1) : m.x, r = k4 < Z4 - 1 ? __ldg(row + 4) : m.w;   // replicate clamp
          const float vals[6] = {l, m.x, m.y, m.z, m.w, r};
          if (a == 1 && b == 1) { ctr[0] = m.x; ctr[1] = m.y; ctr[2] = m.z; ctr[3] = m.w; }
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] += vals[q + c] * ker.w[(a * 3 + b) * 3 + c];
        }
      }
      const unsigned char mq[4] = {mk.x, mk.y, mk.z, mk.w};
      float dq[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        dq[q] = 0.f;
        if (mq[q]) {
          const float e = acc[q] - ctr[q];
          local += e * e;
          dq[q] = -2.f * scale * e;
        }
      }
      d = make_float4(dq[0], dq[1], dq[2], dq[3]);
    }
    *reinterpret_cast<float4*>(dG + comp * V + v) = d;
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
  }
}

__global__ void k_sum_partials(const float* __restrict__ partial, int n, float scale, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) out[0] = t * scale;
  }
}

// scratch: at least vx_smooth_grad_tv_scratch_floats() floats
VX_API int vx_smooth_grad_tv_scratch_floats() { return vx_num_sms() * 16; }

static int smooth_grad_tv_impl(const float* G, const bool* mask, int X, int Y, int Z, const float* weight3_host,
                               float w_over_3n, float* dG, float* scratch, float* loss_out, int skip_unmasked, cudaStream_t st) {
  VxKernel3 ker;
  VX_REQUIRE(fill_kernel(ker, weight3_host, 3) == 0, "vx_smooth_grad_tv", "bad kernel");
  const int64_t n = (int64_t)X * Y * Z * 3;
  if (n <= 0) return 0;
  const bool v4 = vec4_ok(X, Y, Z, 3, G, dG) && (reinterpret_cast<uintptr_t>(mask) & 3) == 0 && (n / 3) % 4 == 0;
  const int blocks = grid_blocks(v4 ? n / 4 : n);
  if (v4) k_smooth_grad_tv_v4<<<blocks, 256, 0, st>>>(G, mask, X, Y, Z, ker, w_over_3n, dG, scratch, skip_unmasked);
  else k_smooth_grad_tv<<<blocks, 256, 0, st>>>(G, mask, X, Y, Z, ker, w_over_3n, dG, scratch);
  int rc = vx_check_launch("vx_smooth_grad_tv");
  if (rc) return rc;
  if (loss_out) {
    k_sum_partials<<<1, 1024, 0, st>>>(scratch, blocks, w_over_3n, loss_out);
    rc = vx_check_launch("vx_smooth_grad_tv(sum)");
  }
  return rc;
}

VX_API int vx_smooth_grad_tv(const float* G, const bool* mask, int X, int Y, int Z, const float* weight3_host,
                             float w_over_3n, float* dG, float* scratch, float* loss_out, cudaStream_t st) {
  return smooth_grad_tv_impl(G, mask, X, Y, Z, weight3_host, w_over_3n, dG, scratch, loss_out, 0, st);
}

// Same, but entries of dG outside the mask are not written (they are zero by definition): for a caller that keeps dG
// zero-initialised and whose mask does not change.  Only the vectorised form (Z % 4 == 0) skips; otherwise identical.
VX_API int vx_smooth_grad_tv_masked_writes(const float* G, const bool* mask, int X, int Y, int Z, const float* weight3_host,
                                           float w_over_3n, float* dG, float* scratch, float* loss_out, cudaStream_t st) {
  return smooth_grad_tv_impl(G, mask, X, Y, Z, weight3_host, w_over_3n, dG, scratch, loss_out, 1, st);
}

// ---------------------------------------------------------------------------------------------
// total_variation_add_grad (total_variation_kernel.cu:14-35; axis-weight quirk kept: the k and the i
// axis both use wz, wx is unused) and total_variation_add_grad_new (:39-66; wx on k, wy on j, wz on i,
// each term times mask[idx]*mask[nb]).  Host divides the weights by 6 (:76-78).
// ---------------------------------------------------------------------------------------------
// One neighbour's contribution: w * clamp(p - q, -1, 1) [* mask_here * mask_there].  The six
// contributions are accumulated in the reference's order (k-, k+, j-, j+, i-, i+) into a local sum
// that is added to grad once, so rounding matches the reference kernel.
template <bool kMasked>
__device__ __forceinline__ float tv_pull(const float* __restrict__ param, const float* __restrict__ mask, size_t here,
                                         size_t there, float w) {
  float t = w * fminf(fmaxf(param[here] - param[there], -1.f), 1.f);
  if (kMasked) t = t * mask[here] * mask[there];
  return t;
}

template <bool kDense, bool kMasked>
__global__ void k_total_variation_add_grad(const float* __restrict__ param, float* __restrict__ grad,
                                           const float* __restrict__ mask, float w_fast, float w_mid, float w_slow,
                                           size_t nx, size_t ny, size_t nz, size_t N) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  if (!kDense && grad[v] == 0) return;  // sparse mode touches only voxels that already have a gradient
  const size_t z = v % nz, y = (v / nz) % ny, x = (v / nz / ny) % nx;  // (channel folds into the leading index)
  const size_t sy = nz, sx = nz * ny;
  float add = 0;
  if (z != 0) add += tv_pull<kMasked>(param, mask, v, v - 1, w_fast);
  if (z != nz - 1) add += tv_pull<kMasked>(param, mask, v, v + 1, w_fast);
  if (y != 0) add += tv_pull<kMasked>(param, mask, v, v - sy, w_mid);
  if (y != ny - 1) add += tv_pull<kMasked>(param, mask, v, v + sy, w_mid);
  if (x != 0) add += tv_pull<kMasked>(param, mask, v, v - sx, w_slow);
  if (x != nx - 1) add += tv_pull<kMasked>(param, mask, v, v + sx, w_slow);
  grad[v] += add;
}

// param/grad: (C, nx, ny, nz) channel-major contiguous, N = numel.  Weight routing follows the
// reference: without a mask the fastest AND the slowest axis use wz and wx is ignored
// (total_variation_kernel.cu:27-32); with a mask it is wx (fastest), wy, wz (slowest) (:53-58).
VX_API int vx_total_variation_add_grad(const float* param, float* grad, const float* mask, float wx, float wy, float wz,
                                       int dense_mode, int64_t sz_i, int64_t sz_j, int64_t sz_k, int64_t N,
                                       cudaStream_t st) {
  if (N <= 0) return 0;
  const int threads = 256;
  const int blocks = (int)((N + threads - 1) / threads);
  wx /= 6; wy /= 6; wz /= 6;  // :76-78
  const float w_fast = mask ? wx : wz, w_mid = wy, w_slow = wz;
  if (!mask && dense_mode && N == sz_i * sz_j * sz_k && vec4_ok((int)sz_i, (int)sz_j, (int)sz_k, 1, param, grad)) {
    k_sdf_reg_backward_v4<false, true><<<grid_blocks(N / 4), 256, 0, st>>>(nullptr, param, (uint32_t)sz_i, (uint32_t)sz_j,
                                                                         (uint32_t)sz_k, 1.f, w_fast, w_mid, w_slow, grad, nullptr);
    return vx_check_launch("vx_total_variation_add_grad");
  }
#define VX_TV(D, M) k_total_variation_add_grad<D, M><<<blocks, threads, 0, st>>>(param, grad, mask, w_fast, w_mid, w_slow, (size_t)sz_i, (size_t)sz_j, (size_t)sz_k, (size_t)N)
  if (mask) { if (dense_mode) VX_TV(true, true); else VX_TV(false, true); }
  else { if (dense_mode) VX_TV(true, false); else VX_TV(false, false); }
#undef VX_TV
  return vx_check_launch("vx_total_variation_add_grad");
}

// vx_fd_gradient_backward followed by the dense, unmasked vx_total_variation_add_grad of a single-channel grid, with one
// read-modify-write of grad (the fine stage's two sdf regularisers, run.py:612-655).  Same result, bit for bit, as the
// two separate calls.
VX_API int vx_sdf_regularisers_backward(const float* dgrad, const float* param, int X, int Y, int Z, float voxel_size,
                                        float wx, float wy, float wz, float* grad, const bool* active, cudaStream_t st) {
  const int64_t V = (int64_t)X * Y * Z;
  if (V <= 0) return 0;
  if (vec4_ok(X, Y, Z, 3, dgrad, grad, param) && (reinterpret_cast<uintptr_t>(active) & 3) == 0) {
    (void)wx;  // the reference's unmasked kernel routes wz to the fastest and the slowest axis (total_variation_kernel.cu:27-32)
    k_sdf_reg_backward_v4<true, true><<<grid_blocks(V / 4), 256, 0, st>>>(dgrad, param, X, Y, Z, voxel_size, wz / 6, wy / 6, wz / 6, grad, active);
    return vx_check_launch("vx_sdf_regularisers_backward");
  }
  int rc = vx_fd_gradient_backward(dgrad, X, Y, Z, voxel_size, grad, st);
  if (rc) return rc;
  return vx_total_variation_add_grad(param, grad, nullptr, wx, wy, wz, 1, X, Y, Z, V, st);
}

// The same two regularisers restricted to the X-slab [x0, x1) of grad (dgrad may be nullptr: TV add-grad only).  dgrad and
// param are full grids: the stencils reach one plane outside the slab.  Used by the slab-sharded data-parallel step
// (reduce-scatter -> TV + Adam on the owned slab -> all-gather of the parameters).
VX_API int vx_sdf_regularisers_backward_slab(const float* dgrad, const float* param, int X, int Y, int Z, float voxel_size,
                                             float wx, float wy, float wz, float* grad, const bool* active, int x0, int x1,
                                             cudaStream_t st) {
  VX_REQUIRE(0 <= x0 && x0 <= x1 && x1 <= X, "vx_sdf_regularisers_backward_slab", "0 <= x0 <= x1 <= X");
  VX_REQUIRE(Z % 4 == 0 && (reinterpret_cast<uintptr_t>(param) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad) & 15) == 0 &&
             (!dgrad || (reinterpret_cast<uintptr_t>(dgrad) & 15) == 0) && (reinterpret_cast<uintptr_t>(active) & 3) == 0 &&
             (int64_t)X * Y * Z < ((int64_t)1 << 32), "vx_sdf_regularisers_backward_slab", "Z % 4 == 0, 16-byte aligned grids, < 2^32 voxels");
  if (x0 == x1) return 0;
  (void)wx;
  const int64_t n4 = (int64_t)(x1 - x0) * Y * (Z / 4);
  if (dgrad)
    k_sdf_reg_backward_v4<true, true><<<grid_blocks(n4), 256, 0, st>>>(dgrad, param, X, Y, Z, voxel_size, wz / 6, wy / 6, wz / 6, grad, active, x0, x1);
  else
    k_sdf_reg_backward_v4<false, true><<<grid_blocks(n4), 256, 0, st>>>(nullptr, param, X, Y, Z, 1.f, wz / 6, wy / 6, wz / 6, grad, nullptr, x0, x1);
  return vx_check_launch("vx_sdf_regularisers_backward_slab");
}

// ---------------------------------------------------------------------------------------------
// Autograd-form total variation (lib/voxurf_fine.py:956-969, used by the coarse stage through
// density_total_variation(sdf_tv) / k0_total_variation, lib/voxurf_coarse.py:300-320):
//   tv = (mean_x + mean_y + mean_z) / 3,  mean_a = mean over neighbour pairs along a with both voxels in the mask
//        of |v[next] - v[here]|
// One pass produces the three pair sums (deterministic two-stage reduction) and d tv / d v in gather form.
// mask: (X,Y,Z) bool shared by all channels, or nullptr.  inv_cnt_host[a] = 1 / (3 * number of masked pairs on a).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sgnf(float d) { return (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f); }

__global__ void k_tv_l1(const float* __restrict__ v, const bool* __restrict__ mask, int C, int X, int Y, int Z,
                        float sx, float sy, float sz, float* __restrict__ grad, float* __restrict__ partial) {
  __shared__ float red[3][32];
  const int64_t V = (int64_t)X * Y * Z;
  float lx = 0.f, ly = 0.f, lz = 0.f;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < V * C; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = t % V;
    const int k = u % Z, j = (u / Z) % Y, i = u / ((int64_t)Z * Y);
    const int64_t sX = (int64_t)Y * Z;
    const bool here = mask ? mask[u] : true;
    float g = 0.f;
    if (here) {
      const float c = __ldg(v + t);
      if (i + 1 < X && (!mask || mask[u + sX])) { const float d = __ldg(v + t + sX) - c; lx += fabsf(d); g -= sgnf(d) * sx; }
      if (i > 0 && (!mask || mask[u - sX])) { g += sgnf(c - __ldg(v + t - sX)) * sx; }
      if (j + 1 < Y && (!mask || mask[u + Z])) { const float d = __ldg(v + t + Z) - c; ly += fabsf(d); g -= sgnf(d) * sy; }
      if (j > 0 && (!mask || mask[u - Z])) { g += sgnf(c - __ldg(v + t - Z)) * sy; }
      if (k + 1 < Z && (!mask || mask[u + 1])) { const float d = __ldg(v + t + 1) - c; lz += fabsf(d); g -= sgnf(d) * sz; }
      if (k > 0 && (!mask || mask[u - 1])) { g += sgnf(c - __ldg(v + t - 1)) * sz; }
    }
    grad[t] = g;
  }
  float l[3] = {lx, ly, lz};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    for (int o = 16; o > 0; o >>= 1) l[a] += __shfl_down_sync(0xffffffffu, l[a], o);
    if ((threadIdx.x & 31) == 0) red[a][threadIdx.x >> 5] = l[a];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float s = (threadIdx.x < (blockDim.x >> 5)) ? red[a][threadIdx.x] : 0.f;
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if (threadIdx.x == 0) partial[a * gridDim.x + blockIdx.x] = s;
    }
  }
}

__global__ void k_tv_l1_finish(const float* __restrict__ partial, int n, float sx, float sy, float sz,
                               float* __restrict__ out) {
  __shared__ float red[32];
  float tot = 0.f;
  for (int a = 0; a < 3; ++a) {
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[a * n + i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
      for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
      if (threadIdx.x == 0) tot += t * (a == 0 ? sx : (a == 1 ? sy : sz));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = tot;
}

// scratch: 3 * vx_smooth_grad_tv_scratch_floats() floats
VX_API int vx_total_variation_l1(const float* v, const bool* mask, int C, int X, int Y, int Z,
                                 const float* inv_cnt_host, float* grad, float* scratch, float* loss_out,
                                 cudaStream_t st) {
  const int64_t n = (int64_t)X * Y * Z * C;
  if (n <= 0) return 0;
  const int blocks = grid_blocks(n);
  k_tv_l1<<<blocks, 256, 0, st>>>(v, mask, C, X, Y, Z, inv_cnt_host[0], inv_cnt_host[1], inv_cnt_host[2], grad, scratch);
  int rc = vx_check_launch("vx_total_variation_l1");
  if (rc) return rc;
  k_tv_l1_finish<<<1, 1024, 0, st>>>(scratch, blocks, inv_cnt_host[0], inv_cnt_host[1], inv_cnt_host[2], loss_out);
  return vx_check_launch("vx_total_variation_l1(sum)");
}
