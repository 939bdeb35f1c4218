/*
 * voxurf_b200.h -- C ABI of libvoxurf_b200.so: the B200-native (sm_100a) operators of Voxurf's
 * ray-batch volume-rendering hot path.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - fp32 data, int64 indices and 1-byte bools at the reference-facing ("legacy") surface,
 *     int32 indices on the fused surface;
 *   - all tensors are contiguous; grids are (C,X,Y,Z) channel-major (channels_last=0, the layout of
 *     the reference's (1,C,X,Y,Z) Parameter) or (X,Y,Z,C) (channels_last=1);
 *   - every call is asynchronous on `stream` (the reference launches on the legacy default stream,
 *     lib/cuda/render_utils_kernel.cu:93; pass the caller's current stream for equivalent ordering);
 *   - return value 0 = ok, otherwise non-zero and vx_last_error() returns a message
 *     (the reference raises through TORCH_CHECK, lib/cuda/render_utils.cpp:46-48);
 *   - n <= 0 inputs are no-ops that return 0 (the reference's early returns, render_utils_kernel.cu:406,465,629).
 *
 * Citations `file:line` are relative to the reference repository (wutong16/Voxurf).
 */
#ifndef VOXURF_B200_H
#define VOXURF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

/* ---- library ------------------------------------------------------------------------------ */
const char* vx_last_error(void);
int vx_abi_version(void);
int vx_sm_count(void);
unsigned long long vx_launch_count(void); /* kernels launched through this library since load */

/* ---- render_utils_cuda (lib/cuda/render_utils.cpp:170-184) --------------------------------- */
/* infer_t_minmax            render_utils.cpp:50-58,  render_utils_kernel.cu:12-35,82-104 */
int vx_infer_t_minmax(const float* rays_o, const float* rays_d, const float* xyz_min, const float* xyz_max,
                      float near, float far, int n_rays, float* t_min, float* t_max, cudaStream_t stream);
/* infer_n_samples           render_utils.cpp:60-65,  render_utils_kernel.cu:38-55,106-121 */
int vx_infer_n_samples(const float* rays_d, const float* t_min, const float* t_max, float stepdist, int n_rays,
                       int64_t* n_samples, cudaStream_t stream);
/* infer_ray_start_dir       render_utils.cpp:67-72,  render_utils_kernel.cu:58-79,123-139 */
int vx_infer_ray_start_dir(const float* rays_o, const float* rays_d, const float* t_min, int n_rays,
                           float* rays_start, float* rays_dir, cudaStream_t stream);
/* sample_pts_on_rays        render_utils.cpp:74-85,  render_utils_kernel.cu:144-242
 * two-phase because the output length is data dependent:
 *   vx_ray_setup  -> t_min, t_max, N_steps, rays_start, rays_dir, offsets[n_rays+1] (exclusive scan, total last)
 *   (caller reads offsets[n_rays], allocates)   vx_sample_fill -> pts, mask_outbbox, ray_id, step_id */
int vx_ray_setup(const float* rays_o, const float* rays_d, const float* xyz_min, const float* xyz_max, float near,
                 float far, float stepdist, int n_rays, float* t_min, float* t_max, int64_t* n_steps,
                 float* rays_start, float* rays_dir, int64_t* offsets, cudaStream_t stream);
int vx_sample_fill(const float* rays_start, const float* rays_dir, const float* xyz_min, const float* xyz_max,
                   const int64_t* offsets, int n_rays, float stepdist, float* rays_pts, bool* mask_outbbox,
                   int64_t* ray_id, int64_t* step_id, cudaStream_t stream);
/* sample_ndc_pts_on_rays    render_utils.cpp:87-97,  render_utils_kernel.cu:245-293 */
int vx_sample_ndc_pts_on_rays(const float* rays_o, const float* rays_d, const float* xyz_min, const float* xyz_max,
                              int n_samples, int n_rays, float* rays_pts, bool* mask_outbbox, cudaStream_t stream);
/* sample_bg_pts_on_rays     render_utils.cpp:99-107, render_utils_kernel.cu:301-360 */
int vx_sample_bg_pts_on_rays(const float* rays_o, const float* rays_d, const float* t_max, float bg_preserve,
                             int n_samples, int n_rays, float* rays_pts, cudaStream_t stream);
/* maskcache_lookup          render_utils.cpp:109-116, render_utils_kernel.cu:367-424 */
int vx_maskcache_lookup(const bool* world, const float* xyz, const float* scale, const float* shift, int sz_i,
                        int sz_j, int sz_k, int64_t n_pts, bool* out, cudaStream_t stream);
/* raw2alpha / raw2alpha_nonuni (interval_vec != NULL)   render_utils.cpp:118-128, kernel.cu:430-504 */
int vx_raw2alpha(const float* density, float shift, const float* interval_vec, float interval, int64_t n,
                 float* exp_d, float* alpha, cudaStream_t stream);
/* raw2alpha_backward / _nonuni_backward                  render_utils.cpp:130-140, kernel.cu:506-574 */
int vx_raw2alpha_backward(const float* exp_d, const float* grad_back, const float* interval_vec, float interval,
                          int64_t n, float* grad, cudaStream_t stream);
/* alpha2weight              render_utils.cpp:142-149, render_utils_kernel.cu:576-651 (bit-exact, warp per ray) */
int vx_alpha2weight(const float* alpha, const int64_t* ray_id, int64_t n_pts, int n_rays, float* weight, float* T,
                    float* alphainv_last, int64_t* i_start, int64_t* i_end, cudaStream_t stream);
/* alpha2weight_backward     render_utils.cpp:151-168, render_utils_kernel.cu:653-707 */
int vx_alpha2weight_backward(const float* alpha, const float* weight, const float* T, const float* alphainv_last,
                             const int64_t* i_start, const int64_t* i_end, int n_rays, int64_t n_pts,
                             const float* grad_weights, const float* grad_last, float* grad, cudaStream_t stream);

/* ---- total_variation_cuda (lib/cuda/total_variation.cpp:29-32) ------------------------------ */
/* total_variation_add_grad (mask == NULL) / total_variation_add_grad_new (float mask)
 * total_variation.cpp:13-27, total_variation_kernel.cu:14-133; in place on grad */
int vx_total_variation_add_grad(const float* param, float* grad, const float* mask, float wx, float wy, float wz,
                                int dense_mode, int64_t sz_i, int64_t sz_j, int64_t sz_k, int64_t numel,
                                cudaStream_t stream);

/* ---- adam_upd_cuda (lib/cuda/adam_upd.cpp:79-86) -------------------------------------------- */
/* mode 0 adam_upd, 1 masked_adam_upd, 2 adam_upd_with_perlr   adam_upd.cpp:36-77, adam_upd_kernel.cu:9-132 */
int vx_adam_upd(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const float* perlr,
                int64_t numel, int step, float beta1, float beta2, float lr, float eps, int mode,
                cudaStream_t stream);
/* the trainer's optimizer: utils.Adam.step / utils.adam, lib/utils.py:83-199 (dense; optional per-voxel lr,
 * optional skip of zero gradients, optional fused zero-fill of grad).  `touched` / `live`: optional bitmaps, one bit
 * per `group` consecutive elements: touched = a gradient was scattered there this step (elsewhere grad == 0 and is
 * not read), live = a gradient was ever scattered there (elsewhere exp_avg == exp_avg_sq == 0: with grad == 0 the
 * dense update is the identity and the element is skipped).  Same result as the dense pass, bit for bit.
 * step_dev (optional, device): {step_size, sqrt_bias_correction2} read by the kernel instead of the by-value arguments,
 * so that a captured CUDA graph can be replayed with new per-step values (likewise inv_s_dev of the fused NeuS kernels). */
int vx_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, const float* perlr, int64_t numel,
                 float beta1, float beta2, float one_minus_beta1, float one_minus_beta2, float step_size,
                 float sqrt_bias_correction2, float eps, int skip_zero_grad, int zero_grad, const uint32_t* touched,
                 const uint32_t* live, int group, const float* step_dev, cudaStream_t stream);
/* vx_adam_step for a single-channel grid with block-level skipping: blocks of 128 consecutive elements whose gradient is
 * entirely zero and that never had a non-zero one (live_blocks[b] == 0, numel / 128 bytes, maintained by the kernel) are
 * the identity under lib/utils.py:154-199 and are not touched; bit-identical to the dense pass */
int vx_adam_step_blocklive(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel, float beta1,
                           float beta2, float one_minus_beta1, float one_minus_beta2, float step_size,
                           float sqrt_bias_correction2, float eps, int zero_grad, uint8_t* live_blocks,
                           const float* step_dev, cudaStream_t stream);
/* The bitmap form of vx_adam_step (lib/utils.py:154-199 restricted to the voxels in touched | live) dealt through a
 * compacted list of the live voxels instead of a static bitmap-word -> warp mapping (the live voxels sit in the surface
 * shell, 1-3 per word: the list keeps every lane busy and every load independent).  group = elements per voxel.
 * work: numel / group + 1 uint32 of scratch.  merge != 0 also performs vx_bitmap_merge in the same pass.
 * Bit-identical to the dense pass. */
int vx_adam_step_worklist(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel, float beta1,
                          float beta2, float one_minus_beta1, float one_minus_beta2, float step_size,
                          float sqrt_bias_correction2, float eps, int zero_grad, uint32_t* touched, uint32_t* live,
                          int group, int merge, uint32_t* work, const float* step_dev, cudaStream_t stream);
/* vx_adam_step_worklist on the slice of a REPLICATED parameter array that this rank owns (data-parallel training, one
 * process per GPU): the kernel also stores every updated parameter into the same slice of the replicas on n_peers other
 * GPUs through NVLink peer memory (peer_params_host: host array of their device addresses) -- optimizer pass and parameter
 * all-gather in one kernel, moving exactly the voxels that changed.  Peer stores are complete when the kernel is: order
 * them against the replicas' readers with a cross-rank barrier on the same stream. */
int vx_adam_step_worklist_peers(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel, float beta1,
                                float beta2, float one_minus_beta1, float one_minus_beta2, float step_size,
                                float sqrt_bias_correction2, float eps, int zero_grad, uint32_t* touched, uint32_t* live,
                                int group, int merge, uint32_t* work, const float* step_dev,
                                const uint64_t* peer_params_host, int n_peers, cudaStream_t stream);
/* vx_adam_step_blocklive on the slab of a replicated single-channel grid (the sdf grid) this rank owns; the parameters of
 * every updated block are also stored into the same slab of the replicas on n_peers other GPUs (peer memory) */
int vx_adam_step_blocklive_peers(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel, float beta1,
                                 float beta2, float one_minus_beta1, float one_minus_beta2, float step_size,
                                 float sqrt_bias_correction2, float eps, int zero_grad, uint8_t* live_blocks,
                                 const float* step_dev, const uint64_t* peer_params_host, int n_peers, cudaStream_t stream);
/* Sparse reduce-scatter of a single-channel gradient grid over peer memory (data-parallel sdf gradient: the ray gradients
 * live in ~10 % of the 128-voxel blocks).  vx_block_nonzero: mask[b] = block b of g has a non-zero element.
 * vx_pull_reduce (after a cross-rank barrier): out = scale * sum, in rank order, of the flagged blocks of n_ranks arrays --
 * grads_host / masks_host: device addresses of the SAME slab (and of its mask bytes) on every rank, own rank included; out
 * may alias the own array; unflagged blocks keep their content.  Bytes moved over NVLink = the flagged blocks only. */
int vx_block_nonzero(const float* g, int64_t numel, uint8_t* mask, cudaStream_t stream);
int vx_pull_reduce(float* out, int64_t numel, const uint64_t* grads_host, const uint64_t* masks_host, int n_ranks,
                   float scale, cudaStream_t stream);
/* CUDA IPC plumbing for the replicas above (one process per GPU, one node): the owner exports the 64-byte handle of a
 * cudaMalloc allocation (its BASE pointer), a peer opens it with its own device current and gets an address its kernels can
 * store to over NVLink; close before the owner frees the allocation. */
int vx_ipc_get_handle(const void* base_ptr, uint8_t* handle_host);
int vx_ipc_open_handle(const uint8_t* handle_host, uint64_t* ptr_out_host);
int vx_ipc_close_handle(uint64_t ptr);
/* cudaDeviceEnablePeerAccess(peer_device) for the current device; 0 also when it was enabled already */
int vx_enable_peer_access(int peer_device);
/* live |= touched; touched = 0 -- after the vx_adam_step that consumed both */
int vx_bitmap_merge(uint32_t* live, uint32_t* touched, int64_t n_words, cudaStream_t stream);

/* ---- torch_scatter.segment_coo(reduce='sum'), sorted index (lib/voxurf_fine.py:753-777) ----- */
int vx_segment_coo_sum(const float* src, const int64_t* index, int64_t M, int K, float* out, cudaStream_t stream);

/* ---- grids: F.grid_sample(bilinear, align_corners=True) call sites --------------------------- */
/* Points are either xyz (P,3) or, when xyz == NULL, implicit (ray_id, step_id) into rays_start/rays_dir:
 * p = start + dir * (stepdist * step), the formula of render_utils_kernel.cu:184-187.
 * The point count is n_host, or *n_dev when n_dev != NULL (sync-free pipelines). */
/* DenseGrid.forward         lib/grid.py:47-58 ; MaskCache gather lib/voxurf_fine.py:936-939 ;
 * coarse grid_sampler       lib/voxurf_coarse.py:435-452 */
int vx_grid_gather(const float* grid, int X, int Y, int Z, int C, int channels_last, const float* xyz_min_host,
                   const float* xyz_max_host, const float* xyz, const int* ray_id, const int* step_id,
                   const float* rays_start, const float* rays_dir, float stepdist, const int* n_dev,
                   int64_t n_host, float* out, cudaStream_t stream);
/* its backward (ATen grid_sampler_3d_backward wrt input): grad_grid += scatter(grad_out) */
int vx_grid_gather_backward(int X, int Y, int Z, int C, int channels_last, const float* xyz_min_host,
                            const float* xyz_max_host, const float* xyz, const int* ray_id, const int* step_id,
                            const float* rays_start, const float* rays_dir, float stepdist, const int* n_dev,
                            int64_t n_host, const float* grad_out, float* grad_grid, uint32_t* touched,
                            cudaStream_t stream);
/* Voxurf.grid_sampler (sample_ret + sample_grad)  lib/voxurf_fine.py:502-534  (L=1, xyz_order=1)
 * Voxurf.sample_sdfs                              lib/voxurf_fine.py:537-577  (L<=8, xyz_order=0) */
int vx_sdf_taps(const float* grid, int X, int Y, int Z, const float* xyz_min_host, const float* xyz_max_host,
                const float* xyz, const int* ray_id, const int* step_id, const float* rays_start,
                const float* rays_dir, float stepdist, const int* n_dev, int64_t n_host,
                const float* displace_host, int L, float voxel_size, int use_grad_norm, int xyz_order,
                float* out_sdf, float* out_feat, float* out_grad, cudaStream_t stream);
int vx_sdf_taps_backward(const float* grid, int X, int Y, int Z, const float* xyz_min_host,
                         const float* xyz_max_host, const float* xyz, const int* ray_id, const int* step_id,
                         const float* rays_start, const float* rays_dir, float stepdist, const int* n_dev,
                         int64_t n_host, const float* displace_host, int L, float voxel_size, int use_grad_norm,
                         int xyz_order, const float* grad_sdf, const float* grad_feat, const float* grad_grad,
                         float* grad_grid, cudaStream_t stream);
/* Mesh field query (lib/voxurf_fine.py:894-910, lib/dvgo_ori.py:679-693): out_u[i][j][k] = (negate ? -1 : 1) * trilinear
 * value of the single-channel grid at (xs[i], ys[j], zs[k]); out_grad (optional, (nx,ny,nz,3), x,y,z order) = the 6-tap
 * gradient of Voxurf.grid_sampler(sample_grad=True), lib/voxurf_fine.py:502-534.  xs may be an X-slab of the lattice axis. */
int vx_sdf_lattice(const float* grid, int X, int Y, int Z, const float* xyz_min_host, const float* xyz_max_host,
                   const float* xs, const float* ys, const float* zs, int nx, int ny, int nz, float voxel_size,
                   int negate, float* out_u, float* out_grad, cudaStream_t stream);
/* neus_alpha_from_sdf_scatter  lib/voxurf_fine.py:463-500 (== lib/voxurf_coarse.py:348-382); give ray_id (int32)
 * or ray_id64; inv_s_dev (optional): 1/s read from device memory instead of the scalar (CUDA-graph replays) */
int vx_neus_alpha(const float* viewdirs, const int* ray_id, const int64_t* ray_id64, const float* sdf,
                  const float* gradient, float dist, float inv_s, const int* n_dev, int64_t n_host, float* alpha,
                  const float* inv_s_dev, cudaStream_t stream);
int vx_neus_alpha_backward(const float* viewdirs, const int* ray_id, const int64_t* ray_id64, const float* sdf,
                           const float* gradient, float dist, float inv_s, const int* n_dev, int64_t n_host,
                           const float* grad_alpha, int accumulate, float* grad_sdf, float* grad_gradient,
                           const float* inv_s_dev, cudaStream_t stream);

/* ---- grid-level stencils ------------------------------------------------------------------- */
/* neus_sdf_gradient('interpolate')  lib/voxurf_fine.py:440-460 ; grad is (3,X,Y,Z); backward accumulates */
int vx_fd_gradient(const float* sdf, int X, int Y, int Z, float voxel_size, float* grad, cudaStream_t stream);
int vx_fd_gradient_backward(const float* dgrad, int X, int Y, int Z, float voxel_size, float* dsdf,
                            cudaStream_t stream);
/* vx_fd_gradient only where `active` (one byte per voxel, groups of 4 along z): the rest of grad is left untouched.  The
 * smooth-gradient TV reads the gradient grid only within one voxel of the non-empty mask (lib/voxurf_fine.py:417-420) */
int vx_fd_gradient_active(const float* sdf, int X, int Y, int Z, float voxel_size, const bool* active, float* grad,
                          cudaStream_t stream);
/* vx_fd_gradient_backward then the dense unmasked vx_total_variation_add_grad(param) with one pass over grad
 * (run.py:612-655: both sdf regularisers of a fine-stage TV iteration); identical result to the two calls.
 * active (optional): voxels outside it have dgrad == 0 at all six neighbours, their FD part is skipped */
int vx_sdf_regularisers_backward(const float* dgrad, const float* param, int X, int Y, int Z, float voxel_size,
                                 float wx, float wy, float wz, float* grad, const bool* active, cudaStream_t stream);
/* the same, restricted to the X-slab [x0, x1) of grad (dgrad may be NULL: TV add-grad only); dgrad / param are full grids.
 * reduce-scatter -> TV + Adam on the owned X-slab -> all-gather of the parameters */
int vx_sdf_regularisers_backward_slab(const float* dgrad, const float* param, int X, int Y, int Z, float voxel_size,
                                      float wx, float wy, float wz, float* grad, const bool* active, int x0, int x1,
                                      cudaStream_t stream);
/* _gaussian_3dconv / tv_smooth_conv: Conv3d(1,1,k,padding=k//2,'replicate')  lib/voxurf_fine.py:236-258 ;
 * B independent volumes; weight_host is (k,k,k) on the HOST, k in {1,3,5} */
int vx_conv3d_replicate(const float* in, int B, int X, int Y, int Z, const float* weight_host, int ksize,
                        float* out, cudaStream_t stream);
int vx_conv3d_replicate_backward(const float* dout, int B, int X, int Y, int Z, const float* weight_host,
                                 int ksize, int accumulate, float* din, cudaStream_t stream);
/* the same convolution (adjoint = 0) or its adjoint (adjoint = 1) for a SEPARABLE kernel w1 (x) w1 (x) w1 -- the reference's
 * normalised Gaussian is one (lib/voxurf_fine.py:246-254) -- as three 1-D passes (15 taps instead of 125 for k = 5);
 * w1_host: the k 1-D weights; scratch: 2 * B*X*Y*Z floats; out (+)= result.  Rounding-level agreement with the k^3 form. */
int vx_conv3d_replicate_separable(const float* in, int B, int X, int Y, int Z, const float* w1_host, int ksize, int adjoint,
                                  int accumulate, float* scratch, float* out, cudaStream_t stream);
/* density_total_variation(smooth_grad_tv)  lib/voxurf_fine.py:417-420: from the FD gradient G (3,X,Y,Z) and the
 * bool nonempty mask: dG = dLoss/dG, loss_out[0] = loss.  w_over_3n = smooth_grad_tv_weight / (3 * mask.sum()) */
int vx_smooth_grad_tv_scratch_floats(void);
int vx_smooth_grad_tv(const float* G, const bool* mask, int X, int Y, int Z, const float* weight3_host,
                      float w_over_3n, float* dG, float* scratch, float* loss_out, cudaStream_t stream);
/* same, but dG is not written outside the mask (zero by definition): for a zero-initialised dG and a static mask */
int vx_smooth_grad_tv_masked_writes(const float* G, const bool* mask, int X, int Y, int Z, const float* weight3_host,
                                    float w_over_3n, float* dG, float* scratch, float* loss_out, cudaStream_t stream);

/* total_variation(v, mask)  lib/voxurf_fine.py:956-969 (autograd form used by the coarse stage): loss_out[0] = tv,
 * grad = d tv / d v.  mask (X,Y,Z) bool shared by the C channels or NULL; inv_cnt_host[a] = 1/(3 * #pairs on axis a);
 * scratch: 3 * vx_smooth_grad_tv_scratch_floats() floats */
int vx_total_variation_l1(const float* v, const bool* mask, int C, int X, int Y, int Z, const float* inv_cnt_host,
                          float* grad, float* scratch, float* loss_out, cudaStream_t stream);

/* Data-parallel k0 exchange, owner side: the all-gathered rows of every rank, recv[r] = [cap x 3 positions | cap x C
 * gradient rows | int32 row count + 3 pad] (what vx_fused_export_k0_rows wrote on rank r), scattered into grad_grid
 * (channels-last, C = 6 or 12) in one launch, restricted to the corners whose voxel has x_lo <= x < x_hi (the X-slab this
 * rank owns; 0, X = the whole grid).  ATen grid_sampler_3d backward arithmetic per corner. */
int vx_k0_rows_scatter(int X, int Y, int Z, int C, const float* xyz_min_host, const float* xyz_max_host, const float* recv,
                       int world, int cap, int x_lo, int x_hi, float* grad_grid, uint32_t* touched, cudaStream_t stream);

/* ---- fused ray march (replaces sample_pts_on_rays + two compactions + MaskCache.forward) ---- */
/* lib/voxurf_fine.py:593-617,631-636,917-942.  bits_* need (offsets[n_rays] >> 5) + n_rays + 1 words. */
int vx_march_flags(const float* rays_start, const float* rays_dir, const float* xyz_min, const float* xyz_max,
                   const int64_t* offsets, int n_rays, float stepdist, const float* mc_density, int mc_X, int mc_Y,
                   int mc_Z, const float* mc_min_host, const float* mc_max_host, float act_shift,
                   float voxel_size_ratio, float thres, uint32_t* bits_inbbox, uint32_t* bits_keep,
                   int* keep_count, int* keep_off /* n_rays+1 */, cudaStream_t stream);
/* Per-cell verdicts of the mask-cache test for vx_march_flags_cells: cells (mc_X, mc_Y, mc_Z) uint8, 1 = every sample whose
 * trilinear cell is (i,j,k) passes `alpha >= thres`, 0 = every one fails, 2 = evaluate exactly (the shell where the mask
 * changes, and the last index of each axis).  Derived from the 8 corner densities of the cell with margins > 10x the fp32
 * evaluation error, so the keep flags of the march are bit-identical with and without the table. */
int vx_mask_cache_cells(const float* mc_density, int mc_X, int mc_Y, int mc_Z, float act_shift, float voxel_size_ratio,
                        float thres, uint8_t* cells, cudaStream_t stream);
/* vx_march_flags with the verdict table (mc_cells may be NULL = vx_march_flags) */
int vx_march_flags_cells(const float* rays_start, const float* rays_dir, const float* xyz_min, const float* xyz_max,
                         const int64_t* offsets, int n_rays, float stepdist, const float* mc_density, int mc_X, int mc_Y,
                         int mc_Z, const float* mc_min_host, const float* mc_max_host, float act_shift,
                         float voxel_size_ratio, float thres, const uint8_t* mc_cells, uint32_t* bits_inbbox,
                         uint32_t* bits_keep, int* keep_count, int* keep_off /* n_rays+1 */, cudaStream_t stream);
/* MaskCache.forward on explicit points  lib/voxurf_fine.py:930-942 */
int vx_mask_cache_query(const float* mc_density, int mc_X, int mc_Y, int mc_Z, const float* mc_min_host,
                        const float* mc_max_host, float act_shift, float voxel_size_ratio, float thres,
                        const float* xyz, int64_t n, bool* out, cudaStream_t stream);
int vx_march_emit(const int64_t* offsets, int n_rays, const uint32_t* bits_keep, const int* keep_off, int capacity,
                  int* ray_id, int* step_id, bool* mask_outbbox /* M0 or NULL */, cudaStream_t stream);
int vx_points_from_steps(const int* ray_id, const int* step_id, const float* rays_start, const float* rays_dir,
                         float stepdist, const int* n_dev, int64_t n_host, float* out, cudaStream_t stream);
/* alpha2weight over int32 per-ray segments with an optional keep flag (alpha > thres filter applied in place,
 * lib/voxurf_fine.py:647-654) and weight > thres flags / per-ray counts (lib/voxurf_fine.py:668-676) */
/* ub360_utils_cuda.cumdist_thres(dist (n_rays, n_pts), thres) -> bool mask (lib/cuda/ub360_utils.cpp:17-25, kernel
 * ub360_utils_kernel.cu:13-33): running distance per ray, restarted where it exceeds thres; bit-identical */
int vx_cumdist_thres(const float* dist, float thres, int n_rays, int n_pts, bool* mask, cudaStream_t stream);
int vx_alpha2weight_seg(const float* alpha, const uint8_t* keep, const int* seg_off, int n_rays, float w_thres,
                        float* weight, float* T, float* alphainv_last, int* i_end, uint8_t* w_keep, int* w_count,
                        cudaStream_t stream);
int vx_alpha2weight_seg_backward(const float* alpha, const float* weight, const float* T, const uint8_t* keep,
                                 const float* alphainv_last, const int* seg_off, const int* i_end, int n_rays,
                                 const float* grad_weights, const float* grad_last, float* grad,
                                 cudaStream_t stream);

/* exclusive scan of n int32 (n ~ rays per batch), out[n] = total */
int vx_scan_i32(const int* in, int n, int* out, cudaStream_t stream);
int vx_sum_f32(const float* x, int n, float* out, cudaStream_t stream);

/* ---- fused, sync-free fine-stage step (every count is read from device memory) -------------------------
 * One Voxurf.forward + loss + backward, lib/voxurf_fine.py:620-802 + run.py:604-639, between the march above and
 * the optimizer: see voxurf_b200/fused.py for the sequence and DESIGN.md for the buffer layout. */
/* grid_sampler(sample_grad) + neus_alpha_from_sdf_scatter + alpha > thres flag   lib/voxurf_fine.py:640-648 */
int vx_fused_sdf_alpha(const float* grid, int X, int Y, int Z, const float* xyz_min_host, const float* xyz_max_host,
                       const int* ray_id, const int* step_id, const float* rays_start, const float* rays_dir,
                       float stepdist, const int* n_dev, const float* viewdirs, float voxel_size, float dist,
                       float inv_s, float thres, float* sdf, float* grad, float* alpha, uint8_t* keep, float* d_w,
                       float* d_sdf_s, float* d_grad_s, const float* inv_s_dev, cudaStream_t stream);
/* weights > thres compaction as an index list                                   lib/voxurf_fine.py:668-676 */
int vx_fused_emit_rows(const uint8_t* w_keep, const int* seg_off, const int* off4, int n_rays, int capacity,
                       int* idx4, int* overflow, cudaStream_t stream);
/* k0 gather + sample_sdfs + positional encodings -> the two MLP input matrices   lib/voxurf_fine.py:678-739 */
int vx_fused_row_features(const float* sdf_grid, const float* k0_grid, int X, int Y, int Z, int C, int k0_channels_last,
                          const float* xyz_min_host, const float* xyz_max_host, const int* ray_id, const int* step_id,
                          const float* rays_start, const float* rays_dir, float stepdist, const int* idx4,
                          const int* n_rows_dev, int capacity, const float* viewdirs, const float* sdf_s,
                          const float* grad_s, float voxel_size, int use_grad_norm, int P, int Vp, int P2, int V2,
                          const float* displace_host, int L, int ld1, int ld2, float* X1, float* X2,
                          cudaStream_t stream);
int vx_fused_fill_logit_cols(const float* logit, int ld_logit, const int* n_rows_dev, int capacity, int col, int ld2,
                             float* X2, cudaStream_t stream);
/* sigmoid + segment_coo compositing + losses + their backward                   lib/voxurf_fine.py:749-763, run.py:604-636 */
int vx_fused_composite_loss(const float* logit1, const float* k_out, int ld_out, const int* idx4, const int* off4,
                            int capacity, const float* weight_s, const float* alphainv_last, const float* target,
                            int n_rays, float w_main, float w_rgb0, float w_ent, float ent_scale, float bg, int train,
                            float* rgb_marched, float* rgb_marched0, float* d_logit1, float* d_kout, float* d_w_s,
                            float* d_last, float* loss_ray, cudaStream_t stream);
/* normal_marched / depth                                                        lib/voxurf_fine.py:765-777 */
int vx_fused_composite_aux(const int* idx4, const int* off4, int capacity, const float* weight_s, const float* grad_s,
                           const int* step_id, float dist, int n_rays, float* normal_marched, float* depth,
                           cudaStream_t stream);
/* backward of the row features into the k0 / sdf gradient grids */
int vx_fused_row_backward(const float* sdf_grid, int X, int Y, int Z, int C, int k0_channels_last,
                          const float* xyz_min_host, const float* xyz_max_host, const int* ray_id, const int* step_id,
                          const float* rays_start, const float* rays_dir, float stepdist, const int* idx4,
                          const int* n_rows_dev, int capacity, float voxel_size, int use_grad_norm, int P, int Vp,
                          int P2, int V2, const float* displace_host, int L, int ld1, int ld2, const float* dX1,
                          const float* dX2, float* d_sdf_s, float* d_grad_s, float* sdf_grad, float* k0_grad,
                          uint32_t* k0_touched, cudaStream_t stream);
/* ---- bit-reproducible gradients: the scatters into 64-bit fixed-point accumulators (integer sums do not depend on the
 * order the atomics land in; SURVEY.md 8e "Determinism" -- the reference inherits ATen's fp32 atomics, run.py:279).
 * acc_scale = accumulator units per 1.0 (2^52: contributions >= 4e-9 keep their fp32 mantissa, |sums| < 2048). */
int vx_fused_row_backward_fx(const float* sdf_grid, int X, int Y, int Z, int C, int k0_channels_last,
                             const float* xyz_min_host, const float* xyz_max_host, const int* ray_id, const int* step_id,
                             const float* rays_start, const float* rays_dir, float stepdist, const int* idx4,
                             const int* n_rows_dev, int capacity, float voxel_size, int use_grad_norm, int P, int Vp,
                             int P2, int V2, const float* displace_host, int L, int ld1, int ld2, const float* dX1,
                             const float* dX2, float* d_sdf_s, float* d_grad_s, int64_t* sdf_acc, int64_t* k0_acc,
                             float acc_scale, uint32_t* k0_touched, cudaStream_t stream);
int vx_fused_alpha_sdf_backward_fx(int X, int Y, int Z, const float* xyz_min_host, const float* xyz_max_host,
                                   const int* ray_id, const int* step_id, const float* rays_start, const float* rays_dir,
                                   float stepdist, const int* n_dev, const float* viewdirs, const float* sdf,
                                   const float* grad, const uint8_t* keep, const float* d_alpha, const float* d_sdf_s,
                                   const float* d_grad_s, float voxel_size, float dist, float inv_s, int64_t* sdf_acc,
                                   float acc_scale, const float* inv_s_dev, cudaStream_t stream);
/* grad[i] += acc[i] / acc_scale and acc[i] = 0 wherever acc[i] != 0; touched (optional): one bit per `group` elements */
int vx_fx_accumulate(int64_t* acc, int64_t n, float acc_scale, float* grad, const uint32_t* touched, int group,
                     cudaStream_t stream);
int vx_mlp_dw_batch_fx(int n_jobs, const int64_t* ptrs_host, const int* dims_host, const int* n_rows_dev, int capacity,
                       int64_t* acc, const float* grad_base, float acc_scale, cudaStream_t stream);
/* data-parallel exchange of the k0 gradient as rows: (xyz, scale * dX2[:, 0:C]) per MLP row, zeros past *n_rows_dev;
 * the gathered rows of all ranks are scattered with vx_grid_gather_backward (n_out, optional: the row count, written
 * next to the rows so that it travels with them).  Pass k0_grad = NULL to vx_fused_row_backward to skip the local scatter. */
int vx_fused_export_k0_rows(const int* ray_id, const int* step_id, const float* rays_start, const float* rays_dir,
                            float stepdist, const int* idx4, const int* n_rows_dev, int capacity, const float* dX2,
                            int ld2, int C, float scale, float* xyz_out, float* g_out, int* n_out, cudaStream_t stream);
/* NeuS-alpha backward + the 7-tap scatter of every M2 sample into the sdf gradient grid */
int vx_fused_alpha_sdf_backward(int X, int Y, int Z, const float* xyz_min_host, const float* xyz_max_host,
                                const int* ray_id, const int* step_id, const float* rays_start, const float* rays_dir,
                                float stepdist, const int* n_dev, const float* viewdirs, const float* sdf,
                                const float* grad, const uint8_t* keep, const float* d_alpha, const float* d_sdf_s,
                                const float* d_grad_s, float voxel_size, float dist, float inv_s, float* sdf_grad,
                                const float* inv_s_dev, cudaStream_t stream);

/* ---- tensor-core MLP (tcgen05, TF32x3 split; lib/voxurf_fine.py:132-187,718,749) ---------------------------- */
/* ---- fused coarse-stage step (lib/voxurf_coarse.py:513-619; csrc/coarse_step.cu) ---------------------------------- */
/* MLP input rows [k0 C | xyz 3 | sin 3P | cos 3P | view 3 | sin 3V | cos 3V | normal 3], normal = gradient / (|gradient| + 1e-5)
 * (lib/voxurf_coarse.py:552-571); rows past *n_rows_dev are zero-filled */
int vx_coarse_row_features(const float* k0_grid, int X, int Y, int Z, int C, int k0_channels_last,
                           const float* xyz_min_host, const float* xyz_max_host, const int* ray_id, const int* step_id,
                           const float* rays_start, const float* rays_dir, float stepdist, const int* idx4,
                           const int* n_rows_dev, int capacity, const float* viewdirs, const float* grad_s, int P, int V,
                           int ld, float* Xrows, cudaStream_t stream);
/* its backward: k0 scatter, and the normal's adjoint ADDED to d_grad_s of each row's sample */
int vx_coarse_row_backward(int X, int Y, int Z, int C, int k0_channels_last, const float* xyz_min_host,
                           const float* xyz_max_host, const int* ray_id, const int* step_id, const float* rays_start,
                           const float* rays_dir, float stepdist, const int* idx4, const int* n_rows_dev, int capacity,
                           const float* grad_s, int P, int V, int ld, const float* dX, float* d_grad_s, float* k0_grad,
                           cudaStream_t stream);
/* rgb_marched = clamp(sum w sigmoid(logit) + (1 - sum w) bg, 0, 1), mse + last-ray entropy, and their backward
 * (lib/voxurf_coarse.py:573-583, run.py:604-610) */
int vx_coarse_composite_loss(const float* logit, int ld_out, const int* idx4, const int* off4, int capacity,
                             const float* weight_s, const float* alphainv_last, const float* target, int n_rays,
                             float w_main, float w_ent, float ent_scale, float bg, int train, float* rgb_marched,
                             float* d_logit, float* d_w_s, float* d_last, float* loss_ray, cudaStream_t stream);
/* y += a * x */
int vx_axpy(const float* x, float a, int64_t n, float* y, cudaStream_t stream);

/* "chunked K-major image" CH(F): IMG[(k/4)*F + f][k%4] -- the UMMA K-major no-swizzle operand layout; a K=32 slice is one
 * contiguous block, fetched with one bulk copy.  vx_mlp_prep writes the hi/lo weight images (k = input feature), zero padded
 * to (Np, Kp), optionally of the transposed matrix (dX chain). */
int vx_mlp_prep(const float* W, int N, int K, int ldw, int Np, int Kp, int transpose, float* W_hi, float* W_lo,
                cudaStream_t stream);
/* up to 16 vx_mlp_prep jobs in one launch: ptrs_host[j*3..] = W, W_hi, W_lo (device addresses);
 * dims_host[j*6..] = N, K, ldw, Np, Kp, transpose */
int vx_mlp_prep_batch(int n_jobs, const int64_t* ptrs_host, const int* dims_host, cudaStream_t stream);
/* Y = chain of up to 4 layers y = act(x W^T + b) [* (mask > 0)] on 128-row tiles.  Packed host arrays (csrc/mlp_tc.cu):
 * ptrs_host[l*5..] = W_hi, W_lo, bias, row image out, ReLU-gate bitmap in (device addresses, 0 = none; see vx_mlp_chain_batch);
 * dims_host[l*4..] = Kp, Np, N, relu.  Row images ACT(F) (raw fp32, r = MLP row, F = feature count padded to 32; byte offset
 * (r/4)*16F + (f/32)*512 + (r%4)*128 + (((f%32)/8) ^ (r%4))*32 + (f%8)*4 -- the MN-major TF32 UMMA operand layout) feed
 * vx_mlp_dw; x_img is the ACT(pad32(Kp[0])) image of the input. */
int vx_mlp_chain(const float* X, int ldx, int K0, const int* n_rows_dev, int capacity, int n_layers,
                 const int64_t* ptrs_host, const int* dims_host, float* Y, int ldy, int n_out, float* x_img,
                 cudaStream_t stream);
/* Up to 2 layer chains ("jobs": the forward chains of both colour networks, or both dX chains) in ONE launch: the 128-row
 * tiles of all jobs share the SMs.  A forward job may end in a tiny last layer evaluated in exact fp32 on the CUDA cores
 * (Wf / bias_f, n_out <= 4 -- lib/voxurf_fine.py:132-149 ends every colour MLP in Linear(width, 3)), and may take some input
 * columns from an earlier job's output rows (k_rgbnet reads rgb_logit.detach(), lib/voxurf_fine.py:741-751).  Per job j:
 *   ptrs_host[j*30 + {0..5}]            = X, x_img, Y, Wf, bias_f, patch           (device addresses, 0 = none)
 *   ptrs_host[j*30 + 6 + l*6 + {0..5}]  = layer l: W_hi, W_lo, bias, row image out, ReLU-gate words in (dX chains: four
 *                                         uint64 per row, feature 64c + 16g + j <-> bit 16c + j of word [4r + g], 1 = pass),
 *                                         ReLU-gate words out (forward chains)
 *   dims_host[j*26 + {0..9}]            = ldx, K0, n_layers, ldy, n_out, ldwf, patch_col, patch_n, patch_ld, dep (-1 = none)
 *   dims_host[j*26 + 10 + l*4 + {0..3}] = layer l: Kp, Np, N, relu
 * done_flags: n_jobs * ceil(capacity / 128) ints (only needed when a job has dep >= 0); zeroed by the call. */
/* development hook: event log of CTA 0 (only in -DMC_TRACE builds of csrc/mlp_tc.cu) */
int vx_mlp_trace_set(int64_t* buf);
int vx_mlp_chain_batch(int n_jobs, const int64_t* ptrs_host, const int* dims_host, const int* n_rows_dev, int capacity,
                       int* done_flags, cudaStream_t stream);
/* split-K weight gradient on the row images: C[m][n] += sum_r A[r][m] B[r][n], c_bias[m] += sum_r A[r][m]
 * (r < *n_rows_dev, m < M_out <= FA, n < N_in <= FB; FA, FB multiples of 32; image rows past *n_rows_dev up to the next
 * multiple of 16 must hold zeros) */
int vx_mlp_dw(const float* A_img, int FA, int M_out, const float* B_img, int FB, int N_in, const int* n_rows_dev,
              int capacity, float* C, int ldc, float* c_bias, cudaStream_t stream);
/* up to 8 vx_mlp_dw GEMMs (same row count) in one launch, the SMs dealt in proportion to each job's cost:
 * ptrs_host[j*4..] = A_img, B_img, C, c_bias (device addresses, c_bias may be 0); dims_host[j*5..] = FA, M_out, FB, N_in, ldc */
int vx_mlp_dw_batch(int n_jobs, const int64_t* ptrs_host, const int* dims_host, const int* n_rows_dev, int capacity,
                    cudaStream_t stream);

/* ---- caller side of the path: view rays and the in-mask-cache ray filter ---------------- */
/* get_rays_of_a_view (lib/voxurf_fine.py:1001-1067): rays_o, rays_d, viewdirs (H*W,3) of one view.  K_host: 3x3 row-major
 * intrinsics, c2w_host: first three rows of the pose (3x4 row-major); mode 0 'lefttop', 1 'center', 2 'random' (jitter_i,
 * jitter_j: (H,W) uniform offsets drawn by the caller); ndc != 0 applies ndc_rays(H, W, K[0][0], ndc_near, ...) */
int vx_rays_of_view(int H, int W, const float* K_host, const float* c2w_host, int inverse_y, int flip_x, int flip_y, int mode,
                    const float* jitter_i, const float* jitter_j, int ndc, float ndc_near, float* rays_o, float* rays_d,
                    float* viewdirs, cudaStream_t stream);
/* hit_coarse_geo (lib/voxurf_fine.py:579-591): hit[r] = some sample of ray r (same sampling as vx_ray_setup /
 * sample_pts_on_rays) is inside the bbox and inside the mask cache; nothing per-sample is materialised */
int vx_rays_hit_mask(const float* rays_o, const float* rays_d, int n_rays, const float* xyz_min_host,
                     const float* xyz_max_host, float near, float far, float stepdist, const float* mc_density, int mc_X,
                     int mc_Y, int mc_Z, const float* mc_min_host, const float* mc_max_host, float act_shift,
                     float voxel_size_ratio, float thres, bool* hit, cudaStream_t stream);
/* the img[mask] / rays[mask] copies of get_training_rays_in_maskcache_sampling (lib/voxurf_fine.py:1150-1156) without a
 * host sync: rows of up to four (n,3) arrays with mask set are appended, in order, at row *top_in of the destinations
 * (rows >= capacity are dropped); *top_out = *top_in + count.  incl = inclusive int32 prefix sum of mask */
int vx_compact_rows3(const bool* mask, const int* incl, int n, const int64_t* top_in, int64_t* top_out, int64_t capacity,
                     const float* src0, const float* src1, const float* src2, const float* src3, float* dst0, float* dst1,
                     float* dst2, float* dst3, cudaStream_t stream);


/* ---- GPU marching cubes (replaces mcubes.marching_cubes(u, threshold), lib/dvgo_ori.py:695-703; csrc/marching.cu) ---- */
/* edge_flag[3 p + a]: the iso-level crosses the edge from lattice point p along axis a; cell_tris[cell]: triangle count */
int vx_mc_classify(const float* u, int nx, int ny, int nz, float thr, const int* tri_count, int* cell_tris,
                   uint8_t* edge_flag, cudaStream_t stream);
/* tri_off / vert_off: inclusive prefix sums of cell_tris / edge_flag; verts (V,3) lattice-index coordinates, tris (T,3) */
int vx_mc_emit(const float* u, int nx, int ny, int nz, float thr, const int* tri_table, const int64_t* tri_off,
               const int* vert_off, const uint8_t* edge_flag, float* verts, int64_t* tris, cudaStream_t stream);

/* development probe: one M = 128, K = 8 TF32 tcgen05.mma on caller-laid-out shared-memory operand images (<= 32 KB each)
 * with caller-chosen descriptor fields; used by tests/test_gpu_mlp.py to pin the operand layouts the kernels rely on */
int vx_umma_probe(const float* A_img, int a_floats, const float* B_img, int b_floats, int64_t desc_a_fields,
                  int64_t desc_b_fields, int64_t idesc, int N, float* D, cudaStream_t stream);

/* development probe: cycles for n_mma back-to-back M = 64 / 128, K = 8 TF32 MMAs of width N (form 0: A from shared memory,
 * 1: A from TMEM; n_acc: 1 accumulator tile or 2 alternating), one value per block */
int vx_umma_rate(int n_blocks, int n_mma, int M, int N, int form, int n_acc, int64_t* cycles, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VOXURF_B200_H */
