/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked or called by the product path.
 *
 * Plain-C restatement of the arithmetic of the reference's native operators
 * (/root/reference/lib/cuda/{render_utils,total_variation,adam_upd}_kernel.cu), one sequential
 * loop per CUDA kernel.  Each function cites the reference lines it follows.
 *
 * Floating-point contract: compile with `-O2 -ffp-contract=off`.  Where nvcc (default
 * -fmad=true) fuses a*b+c of the reference expression into one FMA, this file says fmaf()
 * explicitly; the fusion pattern was read off nvcc 12.9 PTX for the same expressions
 * (DESIGN.md "FMA contraction").  Mixed float/double literals of the reference (1e-6, 1., 1e-3,
 * 1e-10, 1e10) are kept as double literals so the promotions are the same.
 *
 * Pinning status: the reference holds no golden vectors for these kernels.  The
 * restatement is pinned on the GPU box against the reference's own kernels built into
 * oracle/_ref (tests/test_gpu_vs_reference.py); on CPU it is checked against the committed
 * fixtures in tests/golden/.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define VX_EXPORT __attribute__((visibility("default")))

/* ---- sampling: render_utils_kernel.cu:12-35 --------------------------------------------- */
VX_EXPORT void vxo_infer_t_minmax(const float* rays_o, const float* rays_d, const float* xyz_min,
                                  const float* xyz_max, float near, float far, int n_rays,
                                  float* t_min, float* t_max) {
  for (int r = 0; r < n_rays; ++r) {
    const float* o = rays_o + 3 * r;
    const float* d = rays_d + 3 * r;
    /* :23-25  zero component -> 1e-6 (double literal narrowed to float) */
    float v[3], a[3], b[3];
    for (int c = 0; c < 3; ++c) {
      v[c] = (d[c] == 0) ? (float)1e-6 : d[c];
      a[c] = (xyz_max[c] - o[c]) / v[c];  /* :26-28 */
      b[c] = (xyz_min[c] - o[c]) / v[c];  /* :29-31 */
    }
    /* :32-33 */
    t_min[r] = fmaxf(fminf(fmaxf(fmaxf(fminf(a[0], b[0]), fminf(a[1], b[1])), fminf(a[2], b[2])), far), near);
    t_max[r] = fmaxf(fminf(fminf(fminf(fmaxf(a[0], b[0]), fmaxf(a[1], b[1])), fmaxf(a[2], b[2])), far), near);
  }
}

/* nvcc: t = y*y; t = fma(x,x,t); t = fma(z,z,t); sqrt.rn   (render_utils_kernel.cu:48-51,68-71) */
static float ray_norm(const float* d) {
  float t = d[1] * d[1];
  t = fmaf(d[0], d[0], t);
  t = fmaf(d[2], d[2], t);
  return sqrtf(t);
}

/* render_utils_kernel.cu:38-55 */
VX_EXPORT void vxo_infer_n_samples(const float* rays_d, const float* t_min, const float* t_max,
                                   float stepdist, int n_rays, int64_t* n_samples) {
  for (int r = 0; r < n_rays; ++r) {
    const float rnorm = ray_norm(rays_d + 3 * r);
    const float x = ceilf((t_max[r] - t_min[r]) * rnorm / stepdist);
    const double m = fmax((double)x, 1.); /* :53 max(float, 1.) is a double max */
    n_samples[r] = (int64_t)m;
  }
}

/* render_utils_kernel.cu:58-79 */
VX_EXPORT void vxo_infer_ray_start_dir(const float* rays_o, const float* rays_d, const float* t_min,
                                       int n_rays, float* rays_start, float* rays_dir) {
  for (int r = 0; r < n_rays; ++r) {
    const float rnorm = ray_norm(rays_d + 3 * r);
    for (int c = 0; c < 3; ++c) {
      rays_start[3 * r + c] = fmaf(rays_d[3 * r + c], t_min[r], rays_o[3 * r + c]); /* :72-74 */
      rays_dir[3 * r + c] = rays_d[3 * r + c] / rnorm;                             /* :75-77 */
    }
  }
}

/* render_utils_kernel.cu:144-242.  Phase 1 (total_len known from N_steps) is done by the caller;
 * this fills the flat lists.  ray_id/step_id follow :144-164 (scatter-1 + cumsum == "ray r owns
 * N_steps[r] consecutive slots"), points and mask_outbbox follow :167-194. */
VX_EXPORT void vxo_sample_pts_fill(const float* rays_start, const float* rays_dir, const float* xyz_min,
                                   const float* xyz_max, const int64_t* n_steps, int n_rays,
                                   float stepdist, float* rays_pts, uint8_t* mask_outbbox,
                                   int64_t* ray_id, int64_t* step_id) {
  int64_t idx = 0;
  for (int r = 0; r < n_rays; ++r) {
    for (int64_t s = 0; s < n_steps[r]; ++s, ++idx) {
      ray_id[idx] = r;
      step_id[idx] = s;
      const float dist = stepdist * (float)(int)s; /* :184 float * int */
      float p[3];
      for (int c = 0; c < 3; ++c) {
        p[c] = fmaf(rays_dir[3 * r + c], dist, rays_start[3 * r + c]); /* :185-187 */
        rays_pts[3 * idx + c] = p[c];
      }
      /* :191-192 strict comparisons: points on the faces are inside */
      mask_outbbox[idx] = (xyz_min[0] > p[0]) | (xyz_min[1] > p[1]) | (xyz_min[2] > p[2]) |
                          (xyz_max[0] < p[0]) | (xyz_max[1] < p[1]) | (xyz_max[2] < p[2]);
    }
  }
}

/* render_utils_kernel.cu:245-270 */
VX_EXPORT void vxo_sample_ndc_pts(const float* rays_o, const float* rays_d, const float* xyz_min,
                                  const float* xyz_max, int n_samples, int n_rays, float* rays_pts,
                                  uint8_t* mask_outbbox) {
  for (int r = 0; r < n_rays; ++r)
    for (int s = 0; s < n_samples; ++s) {
      const int64_t idx = (int64_t)r * n_samples + s;
      const float dist = ((float)s) / (float)(n_samples - 1); /* :260 */
      float p[3];
      for (int c = 0; c < 3; ++c) {
        p[c] = fmaf(rays_d[3 * r + c], dist, rays_o[3 * r + c]);
        rays_pts[3 * idx + c] = p[c];
      }
      mask_outbbox[idx] = (xyz_min[0] > p[0]) | (xyz_min[1] > p[1]) | (xyz_min[2] > p[2]) |
                          (xyz_max[0] < p[0]) | (xyz_max[1] < p[1]) | (xyz_max[2] < p[2]);
    }
}

/* render_utils_kernel.cu:301-340 (inverted-sphere background samples).  Mixed precision as written:
 * `t_inner - 1. + 1. / (1. - ((float)i_step) / N_samples)` is evaluated in double, narrowed to float. */
VX_EXPORT void vxo_sample_bg_pts(const float* rays_o, const float* rays_d, const float* t_max,
                                 float bg_preserve, int n_samples, int n_rays, float* rays_pts) {
  for (int r = 0; r < n_rays; ++r)
    for (int s = 0; s < n_samples; ++s) {
      const int64_t idx = (int64_t)r * n_samples + s;
      const float t_inner = t_max[r];
      const float ori_t_outer = (float)((double)t_inner - 1. + 1. / (1. - (double)(((float)s) / (float)n_samples)));
      float q[3];
      for (int c = 0; c < 3; ++c) q[c] = fmaf(rays_d[3 * r + c], ori_t_outer, rays_o[3 * r + c]);
      float t = q[1] * q[1];
      t = fmaf(q[0], q[0], t);
      t = fmaf(q[2], q[2], t);
      const float t_outer = sqrtf(t);
      const float m = fmaxf(fabsf(q[0]), fmaxf(fabsf(q[1]), fabsf(q[2])));
      const float R_outer = t_outer / m;
      /* :332 float*float/float in float, then *(1.-bg) and + ... in double */
      const float a = R_outer * R_outer / (t_outer * t_outer);
      const float b = R_outer / t_outer;
      const float o2i = (float)((double)a * (1. - (double)bg_preserve) + (double)(b * bg_preserve));
      for (int c = 0; c < 3; ++c) rays_pts[3 * idx + c] = q[c] * o2i;
    }
}

/* ---- free-space mask: render_utils_kernel.cu:367-424 -------------------------------------- */
VX_EXPORT void vxo_maskcache_lookup(const uint8_t* world, const float* xyz, const float* scale,
                                    const float* shift, int sz_i, int sz_j, int sz_k, int64_t n_pts,
                                    uint8_t* out) {
  memset(out, 0, (size_t)n_pts); /* :405 zeros */
  for (int64_t p = 0; p < n_pts; ++p) {
    /* :385-387 round() = half away from zero, on fma(x, scale, shift) */
    const int i = (int)roundf(fmaf(xyz[3 * p + 0], scale[0], shift[0]));
    const int j = (int)roundf(fmaf(xyz[3 * p + 1], scale[1], shift[1]));
    const int k = (int)roundf(fmaf(xyz[3 * p + 2], scale[2], shift[2]));
    if (0 <= i && i < sz_i && 0 <= j && j < sz_j && 0 <= k && k < sz_k)
      out[p] = world[(int64_t)i * sz_j * sz_k + (int64_t)j * sz_k + k];
  }
}

/* ---- raw2alpha: render_utils_kernel.cu:430-574 ------------------------------------------- */
VX_EXPORT void vxo_raw2alpha(const float* density, float shift, const float* interval_vec,
                             float interval, int64_t n, float* exp_d, float* alpha) {
  for (int64_t i = 0; i < n; ++i) {
    const float iv = interval_vec ? interval_vec[i] : interval;
    const float e = expf(density[i] + shift); /* :439 may be inf */
    exp_d[i] = e;
    alpha[i] = 1 - powf(1 + e, -iv); /* :441 */
  }
}

VX_EXPORT void vxo_raw2alpha_backward(const float* exp_d, const float* grad_back,
                                      const float* interval_vec, float interval, int64_t n,
                                      float* grad) {
  for (int64_t i = 0; i < n; ++i) {
    const float iv = interval_vec ? interval_vec[i] : interval;
    /* :515 min(float, 1e10) is a double min; the product chain is double, narrowed on store */
    const double m = fmin((double)exp_d[i], 1e10);
    grad[i] = (float)(m * (double)powf(1 + exp_d[i], -iv - 1) * (double)iv * (double)grad_back[i]);
  }
}

/* ---- alpha2weight: render_utils_kernel.cu:576-651 ---------------------------------------- */
VX_EXPORT void vxo_alpha2weight(const float* alpha, const int64_t* ray_id, int64_t n_pts, int n_rays,
                                float* weight, float* T, float* alphainv_last, int64_t* i_start,
                                int64_t* i_end) {
  for (int64_t i = 0; i < n_pts; ++i) { weight[i] = 0.f; T[i] = 1.f; } /* :624-625 */
  for (int r = 0; r < n_rays; ++r) { alphainv_last[r] = 1.f; i_start[r] = 0; i_end[r] = 0; } /* :626-628 */
  if (n_pts == 0) return; /* :629 */
  for (int64_t i = 1; i < n_pts; ++i) /* :607-617 */
    if (ray_id[i] != ray_id[i - 1]) { i_start[ray_id[i]] = i; i_end[ray_id[i - 1]] = i; }
  i_end[ray_id[n_pts - 1]] = n_pts; /* :635 */
  for (int r = 0; r < n_rays; ++r) { /* :577-605 */
    const int64_t i_s = i_start[r], i_e_max = i_end[r];
    float T_cum = 1.;
    int64_t i;
    for (i = i_s; i < i_e_max; ++i) {
      T[i] = T_cum;
      weight[i] = T_cum * alpha[i];
      T_cum = (float)((double)T_cum * (1. - (double)alpha[i])); /* :596 double-promoted */
      if ((double)T_cum < 1e-3) { i += 1; break; }               /* :597-600 */
    }
    i_end[r] = i;            /* :602 */
    alphainv_last[r] = T_cum; /* :603 */
  }
}

/* render_utils_kernel.cu:653-707 */
VX_EXPORT void vxo_alpha2weight_backward(const float* alpha, const float* weight, const float* T,
                                         const float* alphainv_last, const int64_t* i_start,
                                         const int64_t* i_end, int n_rays, int64_t n_pts,
                                         const float* grad_weights, const float* grad_last,
                                         float* grad) {
  for (int64_t i = 0; i < n_pts; ++i) grad[i] = 0.f; /* :684 */
  for (int r = 0; r < n_rays; ++r) {
    float back_cum = grad_last[r] * alphainv_last[r]; /* :671 */
    for (int64_t i = i_end[r] - 1; i >= i_start[r]; --i) {
      /* :673  gw*T in float; (1-alpha) in float, +1e-10 and the division in double */
      const float gt = grad_weights[i] * T[i];
      const double den = (double)(1 - alpha[i]) + 1e-10;
      grad[i] = (float)((double)gt - (double)back_cum / den);
      back_cum = fmaf(grad_weights[i], weight[i], back_cum); /* :674 contracted */
    }
  }
}

/* ---- total variation: total_variation_kernel.cu:14-66, host :68-133 ---------------------- */
static float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

/* mask == NULL -> total_variation_add_grad (:14-35; note the k and i axes both use wz, wx unused);
 * mask != NULL -> total_variation_add_grad_new (:39-66; wx on k, wy on j, wz on i). */
VX_EXPORT void vxo_total_variation_add_grad(const float* param, float* grad, const float* mask,
                                            float wx, float wy, float wz, int dense_mode,
                                            int64_t sz_i, int64_t sz_j, int64_t sz_k, int64_t N) {
  wx /= 6; wy /= 6; wz /= 6; /* :76-78 */
  const float wk = mask ? wx : wz, wj = wy, wi = wz;
  const int64_t sj = sz_k, si = sz_k * sz_j;
  for (int64_t x = 0; x < N; ++x) {
    if (!(dense_mode || grad[x] != 0)) continue;
    const int64_t k = x % sz_k, j = x / sz_k % sz_j, i = x / sz_k / sz_j % sz_i;
    float g = 0;
#define TERM(cond, w, nb)                                                                         \
  if (!(cond)) {                                                                                  \
    float t = (w)*clampf(param[x] - param[(nb)], -1.f, 1.f);                                      \
    if (mask) t = t * mask[x] * mask[(nb)];                                                       \
    g += t;                                                                                       \
  }
    TERM(k == 0, wk, x - 1)
    TERM(k == sz_k - 1, wk, x + 1)
    TERM(j == 0, wj, x - sj)
    TERM(j == sz_j - 1, wj, x + sj)
    TERM(i == 0, wi, x - si)
    TERM(i == sz_i - 1, wi, x + si)
#undef TERM
    grad[x] += g;
  }
}

/* ---- Adam: adam_upd_kernel.cu:9-132 ------------------------------------------------------- */
/* mode 0 = adam_upd, 1 = masked_adam_upd (skip grad==0), 2 = adam_upd_with_perlr */
VX_EXPORT void vxo_adam_upd(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                            const float* perlr, int64_t N, int step, float beta1, float beta2,
                            float lr, float eps, int mode) {
  /* host :72  every operand is float, so pow/sqrt resolve to the float overloads */
  const float step_size = lr * sqrtf(1 - powf(beta2, (float)step)) / (1 - powf(beta1, (float)step));
  for (int64_t i = 0; i < N; ++i) {
    if (mode == 1 && grad[i] == 0) continue;
    exp_avg[i] = fmaf(beta1, exp_avg[i], (1 - beta1) * grad[i]);
    exp_avg_sq[i] = fmaf(beta2, exp_avg_sq[i], (1 - beta2) * grad[i] * grad[i]);
    const float s = (mode == 2) ? step_size * perlr[i] : step_size;
    param[i] -= s * exp_avg[i] / (sqrtf(exp_avg_sq[i]) + eps);
  }
}

/* ---- cumdist_thres: ub360_utils_kernel.cu:13-33 -------------------------------------------- */
VX_EXPORT void vxo_cumdist_thres(const float* dist, float thres, int n_rays, int n_pts, unsigned char* mask) {
  for (int r = 0; r < n_rays; ++r) {
    float cum_dist = 0;
    for (int64_t i = (int64_t)r * n_pts; i < (int64_t)(r + 1) * n_pts; ++i) {
      cum_dist += dist[i];
      const int over = (cum_dist > thres);
      cum_dist *= (float)(!over);
      mask[i] = (unsigned char)over;
    }
  }
}
