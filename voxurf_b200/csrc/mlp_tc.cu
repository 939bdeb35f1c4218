// Fused tensor-core MLP for the two colour networks (lib/voxurf_fine.py:132-187,718,749).
//
// fp32-grade accuracy on the 5th-gen tensor cores through a 3-term TF32 split:
//     x = hi + lo,  hi = x with the 13 low mantissa bits cleared (exactly representable in TF32),  lo = x - hi (exact)
//     x*w ~= hi_x*hi_w + hi_x*lo_w + lo_x*hi_w            (missing lo*lo term and TF32 rounding of lo: ~2^-21 relative)
// Three tcgen05.mma.kind::tf32 instructions per K-step accumulate into one fp32 accumulator in TMEM.
//
// One CTA (128 threads) owns a tile of 128 rows and runs the whole layer chain on it:
//   * activations never leave the SM between layers: the epilogue (tcgen05.ld -> +bias -> ReLU) writes the next
//     layer's A operand back on chip: the hi part into TENSOR MEMORY (tcgen05.st; the two hi products run in the
//     TS form, A from TMEM) and the lo part into shared memory in the UMMA K-major no-swizzle layout
//     [K/4][128 rows][4 floats] (one 16-byte chunk per thread per K-chunk: conflict-free stores);
//   * weights (hi/lo images prepared once per optimizer step by k_mlp_prep) are streamed from L2 in K=32 slices by
//     ONE producer thread with bulk async copies (cp.async.bulk + mbarrier complete_tx) into a double-buffered ring that
//     runs ahead across layers and tiles; one thread issues the MMAs and frees ring slots with tcgen05.commit;
//   * the hidden activations are also written to HBM as pre-split hi/lo "chunked K-major images" [row/4][feature][row%4]
//     -- exactly the operand layout of the split-K weight-gradient GEMM (k_mlp_dw), which then needs no staging at all.
// The row count is read from device memory (sync-free pipeline); rows past it are computed as zeros.
#include "common.cuh"

#define MLP_ROWS 128
#define MLP_MAXW 192          // max layer width (N and K)
#define MLP_STAGES 2
#define MLP_SLICE_K 32        // K extent of one weight slice in the cp.async ring (4 MMA K-steps)
#define MLP_MAX_LAYERS 4
#define MLP_TMEM_COLS 512     // D accumulator at column 0, A (hi) operand at column 256
#define MLP_TMEM_A 256
#define MLP_ROW_THREADS 256   // warps 0-7: thread = row (TMEM lane = tid % 128); warps 0-3 / 4-7 take alternate column blocks
#define MLP_THREADS 288       // + warp 8: weight producer

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (warp%4)*32 + t
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// this thread's lane, 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// bulk async copy global -> shared, completion counted on an mbarrier (transaction bytes)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_rows() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 256 row threads only

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
//   core matrix = 8 rows x 16 bytes, rows 16 B apart; SBO = bytes between 8-row groups; LBO = bytes between the two
//   16-byte K-chunks of one MMA.  Our tiles are [K/4][rows][16 B], hence SBO = 128, LBO = rows * 16.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}

// instruction descriptor: D fp32, A/B TF32, both K-major, M = 128
// (both operands K-major: measured on B200, tcgen05.mma.kind::tf32 with an MN-major smem operand accumulates nothing)
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// round-to-nearest TF32 (result has its 13 low mantissa bits clear, so the tensor core reads it exactly)
__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ---------------------------------------------------------------------------------------------
// "Chunked K-major image" CH(F) of a matrix V[k][f] (k = reduction index, f = feature / row of the operand):
//     IMG[(k / 4) * F + f][k % 4]
// i.e. the UMMA K-major no-swizzle operand layout with LBO = F * 16 bytes, SBO = 128 bytes.  A K = 32 slice is a
// contiguous block of 8 * F * 16 bytes, so one bulk copy brings a whole operand slice into shared memory.
// Weights: k = input feature, f = output feature.
//
// Activations / activation gradients go to HBM as ONE raw fp32 "row image" per tensor,
//     ACT[((r / 8) * (F / 4) + f / 4) * 8 + r % 8][f % 4]        (r = MLP row, f = feature)
// chosen for the writer: the epilogue thread (= one MLP row) stores whole float4s and a warp store covers 4 x 128
// contiguous bytes; a 32-row slice is one contiguous block of F * 128 bytes (one bulk copy).  The weight-gradient
// GEMM reduces over r, so its operands must be K-major in r (the tensor core does not take MN-major TF32 operands):
// k_mlp_dw transposes + hi/lo-splits each slice on chip with its otherwise idle threads.
// ---------------------------------------------------------------------------------------------
struct MlpPrepJob {
  const float* W;
  float *W_hi, *W_lo;
  int N, K, ldw, Np, Kp, transpose;
};
#define MLP_PREP_MAX_JOBS 16
struct MlpPrepBatch { MlpPrepJob job[MLP_PREP_MAX_JOBS]; };

// one launch for every weight image of a step: blockIdx.y = job (a layer of a forward or of a transposed dX chain)
__global__ void k_mlp_prep(const __grid_constant__ MlpPrepBatch batch) {
  const MlpPrepJob& J = batch.job[blockIdx.y];
  // logical operand (n, k) = W[n*ldw + k], or W[k*ldw + n] when transposed; zero outside (N, K)
  const int total = J.Np * J.Kp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i & 3, n = (i >> 2) % J.Np, kc = (i >> 2) / J.Np;
    const int k = kc * 4 + j;
    float w = 0.f;
    if (n < J.N && k < J.K) w = J.transpose ? J.W[(int64_t)k * J.ldw + n] : J.W[(int64_t)n * J.ldw + k];
    const float h = tf32_hi(w);
    J.W_hi[i] = h;
    J.W_lo[i] = tf32_hi(w - h);
  }
}

VX_API int vx_mlp_prep_batch(int n_jobs, const int64_t* ptrs_host, const int* dims_host, cudaStream_t st) {
  VX_REQUIRE(n_jobs >= 0 && n_jobs <= MLP_PREP_MAX_JOBS, "vx_mlp_prep_batch", "at most 16 jobs per launch");
  if (n_jobs == 0) return 0;
  MlpPrepBatch b;
  memset(&b, 0, sizeof(b));
  int max_total = 0;
  for (int j = 0; j < n_jobs; ++j) {
    MlpPrepJob& J = b.job[j];
    J.W = reinterpret_cast<const float*>(ptrs_host[3 * j]);
    J.W_hi = reinterpret_cast<float*>(ptrs_host[3 * j + 1]);
    J.W_lo = reinterpret_cast<float*>(ptrs_host[3 * j + 2]);
    J.N = dims_host[6 * j]; J.K = dims_host[6 * j + 1]; J.ldw = dims_host[6 * j + 2];
    J.Np = dims_host[6 * j + 3]; J.Kp = dims_host[6 * j + 4]; J.transpose = dims_host[6 * j + 5];
    VX_REQUIRE(J.W && J.W_hi && J.W_lo, "vx_mlp_prep_batch", "null pointer");
    VX_REQUIRE(J.Np % 16 == 0 && J.Kp % 8 == 0 && J.Np >= J.N && J.Kp >= J.K, "vx_mlp_prep_batch", "bad padding");
    max_total = max(max_total, J.Np * J.Kp);
  }
  k_mlp_prep<<<dim3(vx_blocks(max_total, 256), n_jobs), 256, 0, st>>>(b);
  return vx_check_launch("vx_mlp_prep_batch");
}

VX_API int vx_mlp_prep(const float* W, int N, int K, int ldw, int Np, int Kp, int transpose, float* W_hi, float* W_lo,
                       cudaStream_t st) {
  const int64_t ptrs[3] = {(int64_t)(uintptr_t)W, (int64_t)(uintptr_t)W_hi, (int64_t)(uintptr_t)W_lo};
  const int dims[6] = {N, K, ldw, Np, Kp, transpose};
  return vx_mlp_prep_batch(1, ptrs, dims, st);
}

// ---------------------------------------------------------------------------------------------
// the chain kernel
// ---------------------------------------------------------------------------------------------
struct MlpLayer {
  const float* W_hi;    // CH(Np) image, Kp/4 chunks
  const float* W_lo;
  const float* bias;    // [N] or nullptr
  float* img;           // ACT(Np) row image of this layer's output, or nullptr
  const float* mask;    // ACT(Np) row image of the forward activation: output *= (mask > 0), or nullptr
  int Kp, Np, N, relu;
};
struct MlpChain {
  int n_layers;
  MlpLayer L[MLP_MAX_LAYERS];
};

struct __align__(16) MlpSmem {
  float A_lo[MLP_MAXW / 4 * MLP_ROWS * 4];                           // 96 KB  [K/4][128][4]
  float B[MLP_STAGES][2][MLP_SLICE_K / 4 * MLP_MAXW * 4];            // stages x {hi,lo} x [8 chunks][192][4] = 2 x 48 KB
  float bias[MLP_MAX_LAYERS][MLP_MAXW];
  uint64_t bar_full[MLP_STAGES];
  uint64_t bar_empty[MLP_STAGES];
  uint64_t bar_acc;
  uint32_t tmem_base;
};

// 8 consecutive features [c0, c0+8) of this thread's row: hi -> TMEM (A operand), lo -> smem tile; optional HBM row image
__device__ __forceinline__ void store_a8(MlpSmem& s, uint32_t tmem_a_lane, int row_in_tile, int c0, const float* v,
                                         float* __restrict__ img, int F, int64_t row) {
  float hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { hi[j] = tf32_hi(v[j]); lo[j] = tf32_hi(v[j] - hi[j]); }
  tmem_st8(tmem_a_lane + c0, hi);
#pragma unroll
  for (int q = 0; q < 2; ++q)
    reinterpret_cast<float4*>(s.A_lo)[((c0 >> 2) + q) * MLP_ROWS + row_in_tile] = make_float4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
  if (img) {
    const int64_t q0 = ((row >> 3) * (F >> 2) + (c0 >> 2)) * 8 + (row & 7);   // float4 index of feature quad c0/4
    reinterpret_cast<float4*>(img)[q0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(img)[q0 + 8] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

__global__ void __launch_bounds__(MLP_THREADS, 1)
k_mlp_chain(const float* __restrict__ X, int ldx, int K0, int K0p, const int* __restrict__ n_rows_dev, int capacity, MlpChain ch,
            float* __restrict__ Y, int ldy, int n_out, float* __restrict__ x_img) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  MlpSmem& s = *reinterpret_cast<MlpSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_rows = min(*n_rows_dev, capacity);
  const int n_tiles = (n_rows + MLP_ROWS - 1) / MLP_ROWS;

  if (tid == 0) {
    for (int i = 0; i < MLP_STAGES; ++i) { mbar_init(&s.bar_full[i], 1); mbar_init(&s.bar_empty[i], 1); }
    mbar_init(&s.bar_acc, 1);
    fence_barrier_init();
  }
  for (int l = 0; l < ch.n_layers; ++l)
    for (int c = tid; c < MLP_MAXW; c += MLP_THREADS) s.bias[l][c] = (ch.L[l].bias && c < ch.L[l].N) ? ch.L[l].bias[c] : 0.f;
  if (warp == 0) tmem_alloc(&s.tmem_base, MLP_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s.tmem_base;

  if (warp == MLP_ROW_THREADS / 32) {
    // ===== weight producer: one thread streams every (tile, layer, slice) weight block, two slices ahead at most =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int l = 0; l < ch.n_layers; ++l) {
          const MlpLayer& L = ch.L[l];
          const int NSL = (L.Kp + MLP_SLICE_K - 1) / MLP_SLICE_K;
          for (int sl = 0; sl < NSL; ++sl, ++it) {
            const int slot = it % MLP_STAGES;
            const uint32_t use = it / MLP_STAGES;
            if (use > 0) mbar_wait(&s.bar_empty[slot], (use - 1) & 1);
            const int kchunks = min(MLP_SLICE_K, L.Kp - sl * MLP_SLICE_K) / 4;
            const uint32_t bytes = (uint32_t)kchunks * L.Np * 16;
            mbar_expect_tx(&s.bar_full[slot], 2 * bytes);
            const int64_t off = (int64_t)sl * (MLP_SLICE_K / 4) * L.Np * 4;
            bulk_g2s(&s.B[slot][0][0], L.W_hi + off, bytes, &s.bar_full[slot]);
            bulk_g2s(&s.B[slot][1][0], L.W_lo + off, bytes, &s.bar_full[slot]);
          }
        }
    }
  } else {
    // ===== row threads: stage inputs, (thread 0) issue MMAs, epilogues =====
    const int rt = tid & (MLP_ROWS - 1);        // row within the tile == TMEM lane
    const int half = tid >> 7;                  // 0: even column blocks, 1: odd column blocks
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // a warp may only touch TMEM lanes 32*(warp%4)..
    uint32_t acc_phase = 0, it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int row = tile * MLP_ROWS + rt;
      {
        const bool vec = (ldx % 4 == 0);
        const float* src = X + (int64_t)row * ldx;
        for (int c0 = half * 8; c0 < K0p; c0 += 16) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = 0.f;
          if (row < n_rows) {
            if (vec && c0 + 8 <= K0) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(src + c0));
              const float4 b2 = __ldg(reinterpret_cast<const float4*>(src + c0 + 4));
              v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b2.x; v[5] = b2.y; v[6] = b2.z; v[7] = b2.w;
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (c0 + j < K0) v[j] = __ldg(src + c0 + j);
            }
          }
          store_a8(s, lane_addr + MLP_TMEM_A, rt, c0, v, x_img, K0p, row);
        }
        tmem_st_wait();
      }
      for (int l = 0; l < ch.n_layers; ++l) {
        const MlpLayer& L = ch.L[l];
        const int KS = L.Kp / 8;
        const int NSL = (L.Kp + MLP_SLICE_K - 1) / MLP_SLICE_K;
        const int Np = L.Np;
        fence_proxy_async();   // this thread's A_lo stores -> async proxy
        tc_fence_before();     // this thread's tcgen05.st of the A (hi) operand
        bar_rows();
        if (tid == 0) {
          tc_fence_after();
          const uint32_t idesc = make_idesc_tf32(MLP_ROWS, Np);
          for (int sl = 0; sl < NSL; ++sl) {
            const uint32_t cur = it + sl;
            const int slot = cur % MLP_STAGES;
            mbar_wait(&s.bar_full[slot], (cur / MLP_STAGES) & 1);
            tc_fence_after();
            const int k_steps = min(MLP_SLICE_K / 8, KS - sl * (MLP_SLICE_K / 8));
            for (int kk = 0; kk < k_steps; ++kk) {
              const int ks = sl * (MLP_SLICE_K / 8) + kk;
              const uint32_t a_tm = tmem + MLP_TMEM_A + ks * 8;
              const uint64_t da_lo = make_desc(smem_u32(s.A_lo) + (uint32_t)(ks * 2) * MLP_ROWS * 16, MLP_ROWS * 16, 128);
              const uint32_t b_off = (uint32_t)(kk * 2) * Np * 16;
              const uint64_t db_hi = make_desc(smem_u32(&s.B[slot][0][0]) + b_off, Np * 16, 128);
              const uint64_t db_lo = make_desc(smem_u32(&s.B[slot][1][0]) + b_off, Np * 16, 128);
              umma_tf32_ts(tmem, a_tm, db_hi, idesc, ks > 0);
              umma_tf32_ts(tmem, a_tm, db_lo, idesc, 1);
              umma_tf32_ss(tmem, da_lo, db_hi, idesc, 1);
            }
            umma_commit(&s.bar_empty[slot]);
          }
          umma_commit(&s.bar_acc);
        }
        it += NSL;
        // ---- epilogue: accumulator -> registers -> (+bias, ReLU / mask) -> next A operand (+ HBM images)
        mbar_wait(&s.bar_acc, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        const bool last = (l == ch.n_layers - 1);
        if (!last) {
          for (int c0 = half * 32; c0 < Np; c0 += 64) {
            float v[32];
            tmem_ld32(lane_addr + c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float y = v[j] + s.bias[l][c0 + j];
              if (L.relu) y = fmaxf(y, 0.f);
              v[j] = y;
            }
            if (L.mask && row < n_rows) {
              const float4* mk = reinterpret_cast<const float4*>(L.mask) + (((int64_t)row >> 3) * (Np >> 2) + (c0 >> 2)) * 8 + (row & 7);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 m4 = __ldg(mk + 8 * q);
                if (!(m4.x > 0.f)) v[4 * q] = 0.f;
                if (!(m4.y > 0.f)) v[4 * q + 1] = 0.f;
                if (!(m4.z > 0.f)) v[4 * q + 2] = 0.f;
                if (!(m4.w > 0.f)) v[4 * q + 3] = 0.f;
              }
            }
            if (row >= n_rows) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
              store_a8(s, lane_addr + MLP_TMEM_A, rt, c0 + 8 * q, v + 8 * q, L.img, Np, row);
          }
          tmem_st_wait();
        } else {
          for (int c0 = half * 16; c0 < Np; c0 += 32) {
            float v[16];
            tmem_ld16(lane_addr + c0, v);
            if (row < n_rows) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < n_out) Y[(int64_t)row * ldy + c0 + j] = v[j] + s.bias[l][c0 + j];
            }
          }
          tc_fence_before();
          bar_rows();   // every TMEM read of this tile done before the next tile's first MMA overwrites the accumulator
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, MLP_TMEM_COLS);
}

// Layers are described by two packed HOST arrays so the C ABI stays plain:
//   ptrs_host[l*5 + {0..4}] = device addresses of W_hi, W_lo (CH(Np) images from vx_mlp_prep), bias, row image out,
//                             mask row image (0 = none)
//   dims_host[l*4 + {0..3}] = Kp, Np, N, relu
// X (capacity, ldx) with K0 valid columns; Y (capacity, ldy) receives the first n_out columns of the last layer;
// x_img: optional ACT(K0p) row image of the input.  Every row image needs 128 * ceil(capacity / 128) rows.
VX_API int vx_mlp_chain(const float* X, int ldx, int K0, const int* n_rows_dev, int capacity, int n_layers,
                        const int64_t* ptrs_host, const int* dims_host, float* Y, int ldy, int n_out, float* x_img,
                        cudaStream_t st) {
  VX_REQUIRE(n_layers >= 1 && n_layers <= MLP_MAX_LAYERS, "vx_mlp_chain", "1..4 layers");
  VX_REQUIRE(n_rows_dev != nullptr, "vx_mlp_chain", "n_rows_dev required");
  MlpChain ch;
  ch.n_layers = n_layers;
  for (int l = 0; l < n_layers; ++l) {
    MlpLayer& L = ch.L[l];
    L.W_hi = reinterpret_cast<const float*>(ptrs_host[l * 5 + 0]);
    L.W_lo = reinterpret_cast<const float*>(ptrs_host[l * 5 + 1]);
    L.bias = reinterpret_cast<const float*>(ptrs_host[l * 5 + 2]);
    L.img = reinterpret_cast<float*>(ptrs_host[l * 5 + 3]);
    L.mask = reinterpret_cast<const float*>(ptrs_host[l * 5 + 4]);
    L.Kp = dims_host[l * 4 + 0]; L.Np = dims_host[l * 4 + 1]; L.N = dims_host[l * 4 + 2]; L.relu = dims_host[l * 4 + 3];
    VX_REQUIRE(L.Kp % 8 == 0 && L.Kp >= 8 && L.Kp <= MLP_MAXW && L.Np % 16 == 0 && L.Np >= 16 && L.Np <= MLP_MAXW,
               "vx_mlp_chain", "layer shape");
    if (l + 1 < n_layers)
      VX_REQUIRE(L.Np % 32 == 0 && L.Np == dims_host[(l + 1) * 4 + 0], "vx_mlp_chain", "hidden widths must chain and be multiples of 32");
  }
  const int K0p = dims_host[0];
  VX_REQUIRE(K0 <= K0p && K0 <= ldx && n_out <= dims_host[(n_layers - 1) * 4 + 1], "vx_mlp_chain", "K0 / n_out");
  const int tiles_cap = (capacity + MLP_ROWS - 1) / MLP_ROWS;
  static bool attr_set = false;
  const int smem = (int)sizeof(MlpSmem) + 1024;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { vx_set_error("vx_mlp_chain", cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  if (tiles_cap <= 0) return 0;
  const int blocks = min(tiles_cap, vx_num_sms());
  k_mlp_chain<<<blocks, MLP_THREADS, smem, st>>>(X, ldx, K0, K0p, n_rows_dev, capacity, ch, Y, ldy, n_out, x_img);
  return vx_check_launch("vx_mlp_chain");
}

// ---------------------------------------------------------------------------------------------
// Split-K weight-gradient GEMM on the row images:  C[m][n] += sum_r A[r][m] * B[r][n]   (dW = dY^T H, db = dY^T 1)
// A = ACT(FA) row image of dY (m < M_out <= FA), B = ACT(FB) row image of H (n < N_in <= FB).  Per 32-row slice:
//   1. thread 0 bulk-copies the two raw slices (F * 128 bytes each) into shared memory;
//   2. all 128 threads transpose + hi/lo-split them into four K-major operand tiles (reduction index = MLP row).
//      Reads are consecutive float4s; the scalar writes are bank-conflict free because the operand tiles are padded
//      through the descriptor strides (SBO = 144 B between 8-feature groups, LBO = F/8 * 144 + 32 B between 4-row chunks);
//   3. thread 0 issues the MMAs: both 128-row M tiles (features 0..127 / 128..255) into two TMEM accumulators, plus
//      two N = 16 MMAs against a constant ones tile whose first column gives the bias gradient.
// Slices are dealt round-robin to the CTAs (split-K); partial sums leave as vector atomics.
// ---------------------------------------------------------------------------------------------
#define DW_KC 32
#define DW_THREADS 384                                          // 12 warps transpose; warps 0-3 also run the TMEM epilogue
#define DW_SBO 144                                              // bytes between 8-feature groups of an operand tile
#define DW_TILE_BYTES(F) (8 * (((F) >> 3) * DW_SBO + 32))       // 8 four-row chunks
struct __align__(16) DwSmem {
  float raw[2][DW_KC * MLP_MAXW];                               // bulk-copy landing zone: A slice, B slice (24 KB each)
  uint8_t op[4][DW_TILE_BYTES(MLP_MAXW) + 2048];                // A_hi, A_lo, B_hi, B_lo operand tiles (+ slack: M tile 1 over-read)
  float ones[2 * 16 * 4];                                       // K-major [2 chunks][16 features][4 rows]: feature 0 = 1
  uint64_t bar_full;
  uint64_t bar_mma;
  uint64_t bar_acc;
  uint32_t tmem_base;
};

__device__ __forceinline__ void dw_transpose(const float* __restrict__ raw, int F, uint8_t* __restrict__ op_hi,
                                             uint8_t* __restrict__ op_lo, int tid) {
  const int quads = F >> 2;                       // feature quads per row
  const uint32_t lbo = (uint32_t)(F >> 3) * DW_SBO + 32;
  const int n4 = DW_KC * quads;                   // float4s in the slice
  // float4 idx = ((g * quads + q) * 8 + r8): consecutive threads read consecutive float4s; (g, q) advance incrementally
  const int r8 = tid & 7;
  int q = tid >> 3, g = 0;
  while (q >= quads) { q -= quads; ++g; }
  if (g >= DW_KC / 8) return;
  const uint32_t row_off = (uint32_t)(r8 >> 2) * lbo + (uint32_t)(r8 & 3) * 4;   // chunk parity + row within the 4-row chunk
  for (int idx = tid; idx < n4; idx += DW_THREADS) {
    const float4 x = reinterpret_cast<const float4*>(raw)[idx];
    const int f = 4 * q;
    const uint32_t base = (uint32_t)(2 * g) * lbo + row_off + (uint32_t)(f >> 3) * DW_SBO + (uint32_t)(f & 7) * 16;
    const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float h = tf32_hi(xs[t]);
      *reinterpret_cast<float*>(op_hi + base + t * 16) = h;
      *reinterpret_cast<float*>(op_lo + base + t * 16) = tf32_hi(xs[t] - h);
    }
    q += DW_THREADS / 8;
    while (q >= quads) { q -= quads; ++g; }
  }
}

__global__ void __launch_bounds__(DW_THREADS, 1)
k_mlp_dw(const float* __restrict__ A_img, int FA, int M_out, const float* __restrict__ B_img, int FB, int N_in,
         const int* __restrict__ n_rows_dev, int capacity, float* __restrict__ C, int ldc, float* __restrict__ c_bias) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  DwSmem& s = *reinterpret_cast<DwSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_rows = min(*n_rows_dev, capacity);
  const int n_slices = (n_rows + DW_KC - 1) / DW_KC;
  const int m_tiles = (M_out + MLP_ROWS - 1) / MLP_ROWS;
  if (tid == 0) {
    mbar_init(&s.bar_full, 1); mbar_init(&s.bar_mma, 1); mbar_init(&s.bar_acc, 1);
    fence_barrier_init();
  }
  if (tid < 128) s.ones[tid] = ((tid >> 2) % 16 == 0) ? 1.f : 0.f;   // [chunk][feature][row]: feature 0 of both chunks
  if (warp == 0) tmem_alloc(&s.tmem_base, MLP_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s.tmem_base;
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  const int my_slices = (n_slices > (int)blockIdx.x) ? (n_slices - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const uint32_t a_bytes = DW_KC * FA * 4, b_bytes = DW_KC * FB * 4;
  const uint32_t a_lbo = (uint32_t)(FA >> 3) * DW_SBO + 32, b_lbo = (uint32_t)(FB >> 3) * DW_SBO + 32;
  auto issue_load = [&](int i) {
    const int64_t sl = (int64_t)blockIdx.x + (int64_t)i * gridDim.x;
    mbar_expect_tx(&s.bar_full, a_bytes + b_bytes);
    bulk_g2s(s.raw[0], A_img + sl * DW_KC * FA, a_bytes, &s.bar_full);
    bulk_g2s(s.raw[1], B_img + sl * DW_KC * FB, b_bytes, &s.bar_full);
  };
  if (tid == 0 && my_slices > 0) issue_load(0);
  for (int i = 0; i < my_slices; ++i) {
    mbar_wait(&s.bar_full, i & 1);                           // raw slices of step i landed
    if (i > 0) mbar_wait(&s.bar_mma, (i - 1) & 1);           // MMAs of step i-1 finished reading the operand tiles
    dw_transpose(s.raw[0], FA, s.op[0], s.op[1], tid);
    dw_transpose(s.raw[1], FB, s.op[2], s.op[3], tid);
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      if (i + 1 < my_slices) issue_load(i + 1);              // raw buffers are free again: overlaps with the MMAs below
      tc_fence_after();
      const uint32_t idesc = make_idesc_tf32(MLP_ROWS, FB);
      const uint32_t idesc1 = make_idesc_tf32(MLP_ROWS, 16);
      for (int mt = 0; mt < m_tiles; ++mt) {
#pragma unroll
        for (int kk = 0; kk < DW_KC / 8; ++kk) {             // one MMA K-step = two 4-row chunks
          const uint32_t a_off = (uint32_t)(2 * kk) * a_lbo + (uint32_t)mt * (MLP_ROWS / 8) * DW_SBO;
          const uint32_t b_off = (uint32_t)(2 * kk) * b_lbo;
          const uint64_t da_hi = make_desc(smem_u32(s.op[0]) + a_off, a_lbo, DW_SBO);
          const uint64_t da_lo = make_desc(smem_u32(s.op[1]) + a_off, a_lbo, DW_SBO);
          const uint64_t db_hi = make_desc(smem_u32(s.op[2]) + b_off, b_lbo, DW_SBO);
          const uint64_t db_lo = make_desc(smem_u32(s.op[3]) + b_off, b_lbo, DW_SBO);
          const uint64_t d_ones = make_desc(smem_u32(s.ones), 16 * 16, 128);
          const uint32_t d = tmem + mt * 256;
          const uint32_t acc = (i > 0) || (kk > 0);
          umma_tf32_ss(d, da_hi, db_hi, idesc, acc);
          umma_tf32_ss(d, da_hi, db_lo, idesc, 1);
          umma_tf32_ss(d, da_lo, db_hi, idesc, 1);
          umma_tf32_ss(d + FB, da_hi, d_ones, idesc1, acc);  // bias gradient columns [FB, FB + 16)
          umma_tf32_ss(d + FB, da_lo, d_ones, idesc1, 1);
        }
      }
      umma_commit(&s.bar_mma);
      if (i == my_slices - 1) umma_commit(&s.bar_acc);
    }
  }
  if (my_slices > 0 && warp < 4) {
    mbar_wait(&s.bar_acc, 0);
    tc_fence_after();
    for (int mt = 0; mt < m_tiles; ++mt) {
      const int m = mt * MLP_ROWS + tid;
      for (int c0 = 0; c0 < FB + 16; c0 += 16) {
        float v[16];
        tmem_ld16(lane_addr + mt * 256 + c0, v);   // warp-collective: every thread executes it, only the adds are predicated
        if (m < M_out) {
          if (c0 == FB) {
            if (c_bias) atomicAdd(c_bias + m, v[0]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int n = c0 + 4 * q;
              if (n + 3 < N_in && (ldc % 4 == 0)) {
                atomicAdd(reinterpret_cast<float4*>(C + (int64_t)m * ldc + n), make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (n + j < N_in) atomicAdd(C + (int64_t)m * ldc + n + j, v[4 * q + j]);
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, MLP_TMEM_COLS);
}

VX_API int vx_mlp_dw(const float* A_img, int FA, int M_out, const float* B_img, int FB, int N_in, const int* n_rows_dev,
                     int capacity, float* C, int ldc, float* c_bias, cudaStream_t st) {
  VX_REQUIRE(n_rows_dev != nullptr, "vx_mlp_dw", "n_rows_dev required");
  VX_REQUIRE(FA % 8 == 0 && FA >= 8 && FA <= MLP_MAXW && FB % 16 == 0 && FB >= 16 && FB <= MLP_MAXW && M_out >= 1 &&
             M_out <= FA && N_in >= 1 && N_in <= FB, "vx_mlp_dw", "shape");
  static bool attr_set = false;
  const int smem = (int)sizeof(DwSmem) + 1024;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { vx_set_error("vx_mlp_dw", cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  const int slices_cap = (capacity + DW_KC - 1) / DW_KC;
  if (slices_cap <= 0) return 0;
  const int gx = max(1, min(vx_num_sms(), (slices_cap + 3) / 4));
  k_mlp_dw<<<gx, DW_THREADS, smem, st>>>(A_img, FA, M_out, B_img, FB, N_in, n_rows_dev, capacity, C, ldc, c_bias);
  return vx_check_launch("vx_mlp_dw");
}
