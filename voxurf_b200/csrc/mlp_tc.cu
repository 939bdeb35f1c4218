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
//   * the activations (and, in the dX chain, their gradients) are also written to HBM as raw fp32 "row images" in the
//     MN-major 32-byte-swizzled UMMA operand layout, which the split-K weight-gradient GEMM (k_mlp_dw) reads with plain
//     bulk copies: no transposition anywhere.
// The row count is read from device memory (sync-free pipeline); rows past it are computed as zeros.
#include "common.cuh"

#define MLP_ROWS 128
#define MLP_MAXW 192          // max layer width (N and K)
#define MLP_STAGES 2
#define MLP_SLICE_K 32        // K extent of one weight slice in the bulk-copy ring (4 MMA K-steps); a 5 x K=16 ring measured 10 % slower
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

// MN-major TF32 operand: layout type 1 (128-byte swizzle, 32-byte base) -- with any other layout type an MN-major TF32
// MMA accumulates nothing (measured).  LBO = bytes between 32-element MN groups, SBO = bytes between 4-element K groups.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return make_desc(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)1 << 61);
}

// instruction descriptor: D fp32, A/B TF32, both K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// round-to-nearest TF32 (result has its 13 low mantissa bits clear, so the tensor core reads it exactly)
__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// the same rounding (nearest, ties away from zero) in two integer instructions, for finite inputs: the cvt above is
// expanded by ptxas into add + mask + an Inf/NaN guard, which the hot loops below do not need
__device__ __forceinline__ float tf32_rn(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// ---------------------------------------------------------------------------------------------
// "Chunked K-major image" CH(F) of a matrix V[k][f] (k = reduction index, f = feature / row of the operand):
//     IMG[(k / 4) * F + f][k % 4]
// i.e. the UMMA K-major no-swizzle operand layout with LBO = F * 16 bytes, SBO = 128 bytes.  A K = 32 slice is a
// contiguous block of 8 * F * 16 bytes, so one bulk copy brings a whole operand slice into shared memory.
// Weights: k = input feature, f = output feature.
//
// Activations / activation gradients go to HBM as ONE raw fp32 "row image" ACT(F) per tensor (r = MLP row, f = feature,
// F a multiple of 32), byte offset
//     (r / 4) * 16 F  +  (f / 32) * 512  +  (r % 4) * 128  +  (((f % 32) / 8) ^ (r % 4)) * 32  +  (f % 8) * 4
// i.e. 4-row x 32-feature atoms of 512 bytes; inside an atom every row is one 128-byte line whose four 32-byte
// pieces are XOR-permuted with the row number.  This is exactly what tcgen05.mma.kind::tf32 reads as an MN-major
// operand (shared-memory descriptor layout type 1, "128B swizzle with 32B base", the only MN-major layout TF32 has;
// LBO = 512 between 32-feature groups, SBO = 16 F between 4-row groups; measured with vx_umma_probe,
// scripts/probe_umma.py): the weight-gradient GEMM reduces over r, so for it a slice of 16 rows (one contiguous block
// of 64 F bytes, one bulk copy) is a ready-made operand.  It also suits the writer: an epilogue thread (= one MLP
// row) stores 8 consecutive features as one aligned 32-byte piece.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int64_t act_offset(int64_t r, int f, int F) {   // in floats
  return ((r >> 2) * (F >> 5) + (f >> 5)) * 128 + (r & 3) * 32 + ((((f & 31) >> 3) ^ (int)(r & 3)) << 3) + (f & 7);
}

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
  for (int j = 0; j < 8; ++j) { hi[j] = tf32_rn(v[j]); lo[j] = tf32_rn(v[j] - hi[j]); }
  tmem_st8(tmem_a_lane + c0, hi);
#pragma unroll
  for (int q = 0; q < 2; ++q)
    reinterpret_cast<float4*>(s.A_lo)[((c0 >> 2) + q) * MLP_ROWS + row_in_tile] = make_float4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
  if (img) {   // one aligned 32-byte piece, one 256-bit store
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(img + act_offset(row, c0, F)), "f"(v[0]),
                 "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
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
        // this thread's share of the row, six 8-float pieces at a time: all loads are issued before the first use, so
        // the tile pays one memory latency instead of one per piece
        for (int cbase = half * 8; cbase < K0p; cbase += 16 * 6) {
          float v[6][8];
#pragma unroll
          for (int u = 0; u < 6; ++u) {
            const int c0 = cbase + 16 * u;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
            if (c0 < K0p && row < n_rows) {
              if (vec && c0 + 8 <= K0) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(src + c0));
                const float4 b2 = __ldg(reinterpret_cast<const float4*>(src + c0 + 4));
                v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w;
                v[u][4] = b2.x; v[u][5] = b2.y; v[u][6] = b2.z; v[u][7] = b2.w;
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (c0 + j < K0) v[u][j] = __ldg(src + c0 + j);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 6; ++u) {
            const int c0 = cbase + 16 * u;
            if (c0 < K0p) store_a8(s, lane_addr + MLP_TMEM_A, rt, c0, v[u], x_img, (K0p + 31) & ~31, row);
          }
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
            float v[32], mk[32];
            const bool gate = L.mask && row < n_rows;
            if (gate) {
              // ReLU gates of the 32 features [c0, c0 + 32) of this row: one 128-byte line of the forward row image
              // (32-byte pieces permuted), four 256-bit loads issued ahead of the TMEM load they are applied to
              const float* line = L.mask + act_offset(row, c0, Np) - ((((c0 & 31) >> 3) ^ (row & 3)) << 3);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                             : "=f"(mk[8 * q]), "=f"(mk[8 * q + 1]), "=f"(mk[8 * q + 2]), "=f"(mk[8 * q + 3]), "=f"(mk[8 * q + 4]),
                               "=f"(mk[8 * q + 5]), "=f"(mk[8 * q + 6]), "=f"(mk[8 * q + 7])
                             : "l"(line + ((q ^ (row & 3)) << 3)));
            }
            tmem_ld32(lane_addr + c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float y = v[j] + s.bias[l][c0 + j];
              if (L.relu) y = fmaxf(y, 0.f);
              v[j] = y;
            }
            if (gate) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (!(mk[j] > 0.f)) v[j] = 0.f;
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
// Split-K weight-gradient GEMMs on the row images:  C[m][n] += sum_r A[r][m] * B[r][n]   (dW = dY^T H, db = dY^T 1)
// A = ACT(FA) row image of dY (m < M_out <= FA), B = ACT(FB) row image of H (n < N_in <= FB).
// One launch runs up to 8 such GEMMs (all layers of both colour networks): the CTAs are dealt to the jobs in
// proportion to their cost and every CTA walks its job's 16-row slices round-robin through a three-role pipeline
//   warp 13 lane 0  producer : bulk-copies the raw A / B slices (64 F bytes each) into a 6-deep ring;
//   warps 0-11      split a slice element-wise (hi in place, lo into a twin buffer; 128-bit loads / stores) -- the row
//                   image already is the MN-major operand layout, nothing is transposed;
//   warp 12 lane 0  issues the MMAs of a slice, both operands MN-major (instruction-descriptor bits 15 / 16): both
//                   128-row M tiles (features 0..127 / 128..255) into two TMEM accumulators, N = FB + 16 wide: the B slice
//                   carries a constant ones feature whose accumulator column is the bias gradient; tcgen05.commit hands
//                   the stage back to the producer.
// Partial sums leave as vector atomics (warps 0-3).
// ---------------------------------------------------------------------------------------------
#define DW_KC 16
#define DW_T_THREADS 384                                        // 12 splitting warps; warps 0-3 also run the TMEM epilogue
#define DW_THREADS (DW_T_THREADS + 64)                          // + MMA warp + producer warp
#define DW_MAX_JOBS 8
struct DwJob {
  const float* A;
  const float* B;
  float* C;
  float* c_bias;
  int FA, M_out, FB, N_in, ldc;
  int cta_begin, cta_count;
};
struct DwBatch {
  int n_jobs;
  DwJob job[DW_MAX_JOBS];
};
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

#define DW_RAW_STAGES 6      // raw slices in flight (hi parts are produced in place): covers the HBM latency
#define DW_LO_STAGES 2       // lo twins only live from the split to the MMAs of a slice
// The B slice lives in shared memory with one extra 32-feature group per 4-row group (SBO = (FB/32 + 1) * 512 bytes):
// feature FB of the hi part is the constant 1 (everything else in the group and the whole group of the lo part is 0),
// written once at kernel start and never touched by the copies.  The MMAs then run with N = FB + 16 and column FB of
// the accumulator is sum_r A[r][m], the bias gradient, for free -- a separate N = 16 MMA would cost the 96-clock floor of
// every tcgen05.mma, as much as N = 128.
#define DW_A_FLOATS (DW_KC * MLP_MAXW)                          // 12 KB
#define DW_B_FLOATS ((DW_KC / 4) * (MLP_MAXW / 32 + 1) * 128)   // 14 KB
struct __align__(16) DwmSmem {
  float rawA[DW_RAW_STAGES][DW_A_FLOATS];                       // A slices (raw -> hi)
  float rawB[DW_RAW_STAGES][DW_B_FLOATS];                       // B slices (raw -> hi) + ones group
  float loA[DW_LO_STAGES][DW_A_FLOATS];
  float loB[DW_LO_STAGES][DW_B_FLOATS];
  float slack[1024];                                            // M tile 1 over-reads up to 2 KB past an A slice
  uint64_t full[DW_RAW_STAGES], empty[DW_RAW_STAGES];
  uint64_t split[DW_LO_STAGES], lo_empty[DW_LO_STAGES];
  uint64_t bar_acc;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t make_idesc_tf32_major(int M, int N, int a_mn, int b_mn) {
  return make_idesc_tf32(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}

__device__ __forceinline__ void dw_split4(float4* __restrict__ hi, float4* __restrict__ lo, int idx, const float4 x) {
  const float4 h = make_float4(tf32_rn(x.x), tf32_rn(x.y), tf32_rn(x.z), tf32_rn(x.w));
  hi[idx] = h;
  lo[idx] = make_float4(tf32_rn(x.x - h.x), tf32_rn(x.y - h.y), tf32_rn(x.z - h.z), tf32_rn(x.w - h.w));
}

__global__ void __launch_bounds__(DW_THREADS, 1)
k_mlp_dw(const __grid_constant__ DwBatch batch, const int* __restrict__ n_rows_dev, int capacity) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  DwmSmem& s = *reinterpret_cast<DwmSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int ji = 0;
  while (ji + 1 < batch.n_jobs && (int)blockIdx.x >= batch.job[ji].cta_begin + batch.job[ji].cta_count) ++ji;
  const DwJob& J = batch.job[ji];
  const int cta = (int)blockIdx.x - J.cta_begin, n_cta = J.cta_count;
  const int FA = J.FA, FB = J.FB;
  const int n_rows = min(*n_rows_dev, capacity);
  const int n_slices = (n_rows + DW_KC - 1) / DW_KC;
  const int m_tiles = (J.M_out + MLP_ROWS - 1) / MLP_ROWS;
  if (tid == 0) {
    for (int i = 0; i < DW_RAW_STAGES; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
    for (int i = 0; i < DW_LO_STAGES; ++i) { mbar_init(&s.split[i], DW_T_THREADS / 32); mbar_init(&s.lo_empty[i], 1); }
    mbar_init(&s.bar_acc, 1);
    fence_barrier_init();
  }
  const int b_groups = FB >> 5;                             // 32-feature groups of the B image; the ones group follows
  const uint32_t b_sbo = (uint32_t)(b_groups + 1) * 512;    // bytes between 4-row groups of the B slice in shared memory
  for (int i = tid; i < (DW_RAW_STAGES + DW_LO_STAGES) * (DW_KC / 4) * 128; i += DW_THREADS) {
    const int buf = i / ((DW_KC / 4) * 128), kg = (i / 128) % (DW_KC / 4), e = i % 128;   // e: float within the 512-byte atom
    float* base = (buf < DW_RAW_STAGES ? s.rawB[buf] : s.loB[buf - DW_RAW_STAGES]) + kg * (b_sbo / 4) + b_groups * 128;
    // feature 0 of the group, row r of the 4-row group: byte (r * 128) + ((0 ^ r) * 32)
    base[e] = (buf < DW_RAW_STAGES && (e & 31) == ((e >> 5) << 3)) ? 1.f : 0.f;
  }
  if (warp == 0) tmem_alloc(&s.tmem_base, MLP_TMEM_COLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s.tmem_base;
  const int my_slices = (n_slices > cta) ? (n_slices - 1 - cta) / n_cta + 1 : 0;
  const uint32_t a_bytes = DW_KC * FA * 4, b_bytes = DW_KC * FB * 4;

  if (warp == DW_T_THREADS / 32 + 1) {
    // ===== producer =====
    if (lane == 0)
      for (int i = 0; i < my_slices; ++i) {
        const int st = i % DW_RAW_STAGES;
        if (i >= DW_RAW_STAGES) mbar_wait(&s.empty[st], (i / DW_RAW_STAGES - 1) & 1);
        const int64_t sl = (int64_t)cta + (int64_t)i * n_cta;
        mbar_expect_tx(&s.full[st], a_bytes + b_bytes);
        bulk_g2s(s.rawA[st], J.A + sl * DW_KC * FA, a_bytes, &s.full[st]);
#pragma unroll
        for (int kg = 0; kg < DW_KC / 4; ++kg)   // one copy per 4-row group: the shared-memory stride leaves room for the ones group
          bulk_g2s(s.rawB[st] + kg * (b_sbo / 4), J.B + sl * DW_KC * FB + kg * 4 * FB, b_bytes / (DW_KC / 4), &s.full[st]);
      }
  } else if (warp == DW_T_THREADS / 32) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32_major(MLP_ROWS, FB + 16, 1, 1);   // + the ones group: column FB = bias gradient
      const uint32_t a_sbo = 16u * FA;                             // bytes between 4-row groups; 512 between 32-feature groups
      // descriptors of stage 0 / K-step 0 / M tile 0; the others differ only in the start-address field (bytes >> 4)
      const uint64_t da_hi0 = make_desc_mn(smem_u32(s.rawA[0]), 512, a_sbo), db_hi0 = make_desc_mn(smem_u32(s.rawB[0]), 512, b_sbo);
      const uint64_t da_lo0 = make_desc_mn(smem_u32(s.loA[0]), 512, a_sbo), db_lo0 = make_desc_mn(smem_u32(s.loB[0]), 512, b_sbo);
      constexpr uint32_t kStageA = DW_A_FLOATS * 4 >> 4, kStageB = DW_B_FLOATS * 4 >> 4;
      for (int i = 0; i < my_slices; ++i) {
        const int rs = i % DW_RAW_STAGES, ls = i % DW_LO_STAGES;
        mbar_wait(&s.split[ls], (i / DW_LO_STAGES) & 1);
        tc_fence_after();
        for (int mt = 0; mt < m_tiles; ++mt) {
#pragma unroll
          for (int kk = 0; kk < DW_KC / 8; ++kk) {             // one MMA K-step = two 4-row groups
            const uint32_t a_off = ((uint32_t)(2 * kk) * a_sbo + (uint32_t)mt * (MLP_ROWS / 32) * 512) >> 4;
            const uint32_t b_off = ((uint32_t)(2 * kk) * b_sbo) >> 4;
            const uint64_t da_hi = da_hi0 + rs * kStageA + a_off, da_lo = da_lo0 + ls * kStageA + a_off;
            const uint64_t db_hi = db_hi0 + rs * kStageB + b_off, db_lo = db_lo0 + ls * kStageB + b_off;
            const uint32_t d = tmem + mt * 256;
            const uint32_t acc = (i > 0) || (kk > 0);
            umma_tf32_ss(d, da_hi, db_hi, idesc, acc);
            umma_tf32_ss(d, da_hi, db_lo, idesc, 1);
            umma_tf32_ss(d, da_lo, db_hi, idesc, 1);
          }
        }
        umma_commit(&s.empty[rs]);
        umma_commit(&s.lo_empty[ls]);
      }
      if (my_slices > 0) umma_commit(&s.bar_acc);
    }
  } else {
    // ===== splitters =====
    const int na4 = DW_KC * FA / 4, nb4 = DW_KC * FB / 4;
    for (int i = 0; i < my_slices; ++i) {
      const int rs = i % DW_RAW_STAGES, ls = i % DW_LO_STAGES;
      mbar_wait(&s.full[rs], (i / DW_RAW_STAGES) & 1);
      if (i >= DW_LO_STAGES) mbar_wait(&s.lo_empty[ls], (i / DW_LO_STAGES - 1) & 1);
      float4* hiA = reinterpret_cast<float4*>(s.rawA[rs]);
      float4* hiB = reinterpret_cast<float4*>(s.rawB[rs]);
      float4* loA = reinterpret_cast<float4*>(s.loA[ls]);
      float4* loB = reinterpret_cast<float4*>(s.loB[ls]);
      const int b_sbo4 = (int)(b_sbo >> 4);                        // float4s per 4-row group of the B slice in shared memory
      if (na4 == 2 * DW_T_THREADS && nb4 == 2 * DW_T_THREADS) {   // 192 x 192: four independent 128-bit loads per thread
        const int i0 = tid, i1 = tid + DW_T_THREADS;
        const int j0 = (i0 / MLP_MAXW) * b_sbo4 + i0 % MLP_MAXW, j1 = (i1 / MLP_MAXW) * b_sbo4 + i1 % MLP_MAXW;
        const float4 a0 = hiA[i0], a1 = hiA[i1], b0 = hiB[j0], b1 = hiB[j1];
        dw_split4(hiA, loA, i0, a0);
        dw_split4(hiA, loA, i1, a1);
        dw_split4(hiB, loB, j0, b0);
        dw_split4(hiB, loB, j1, b1);
      } else {
        for (int idx = tid; idx < max(na4, nb4); idx += DW_T_THREADS) {
          const bool va = idx < na4, vb = idx < nb4;
          const int j = (idx / FB) * b_sbo4 + idx % FB;             // FB float4s per 4-row group
          float4 xa, xb;
          if (va) xa = hiA[idx];
          if (vb) xb = hiB[j];
          if (va) dw_split4(hiA, loA, idx, xa);
          if (vb) dw_split4(hiB, loB, j, xb);
        }
      }
      fence_proxy_async();           // every thread: its own stores -> async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.split[ls]);   // one arrival per warp (384 single arrivals on one word serialise)
    }
    if (my_slices > 0 && warp < 4) {
      const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
      mbar_wait(&s.bar_acc, 0);
      tc_fence_after();
      for (int mt = 0; mt < m_tiles; ++mt) {
        const int m = mt * MLP_ROWS + tid;
        for (int c0 = 0; c0 < FB + 16; c0 += 16) {
          float v[16];
          tmem_ld16(lane_addr + mt * 256 + c0, v);   // warp-collective: every thread executes it, only the adds are predicated
          if (m < J.M_out) {
            if (c0 == FB) {
              if (J.c_bias) atomicAdd(J.c_bias + m, v[0]);
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int n = c0 + 4 * q;
                if (n + 3 < J.N_in && (J.ldc % 4 == 0)) {
                  atomicAdd(reinterpret_cast<float4*>(J.C + (int64_t)m * J.ldc + n), make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
                } else {
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    if (n + j < J.N_in) atomicAdd(J.C + (int64_t)m * J.ldc + n + j, v[4 * q + j]);
                }
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

// ptrs_host[j*4..] = A_img, B_img, C, c_bias (device addresses; c_bias may be 0); dims_host[j*5..] = FA, M_out, FB, N_in, ldc
VX_API int vx_mlp_dw_batch(int n_jobs, const int64_t* ptrs_host, const int* dims_host, const int* n_rows_dev, int capacity,
                           cudaStream_t st) {
  VX_REQUIRE(n_rows_dev != nullptr, "vx_mlp_dw_batch", "n_rows_dev required");
  VX_REQUIRE(n_jobs >= 0 && n_jobs <= DW_MAX_JOBS, "vx_mlp_dw_batch", "at most 8 jobs per launch");
  const int slices_cap = (capacity + DW_KC - 1) / DW_KC;
  if (n_jobs == 0 || slices_cap <= 0) return 0;
  DwBatch b;
  memset(&b, 0, sizeof(b));
  b.n_jobs = n_jobs;
  double cost[DW_MAX_JOBS], total = 0;
  for (int j = 0; j < n_jobs; ++j) {
    DwJob& J = b.job[j];
    J.A = reinterpret_cast<const float*>(ptrs_host[4 * j]);
    J.B = reinterpret_cast<const float*>(ptrs_host[4 * j + 1]);
    J.C = reinterpret_cast<float*>(ptrs_host[4 * j + 2]);
    J.c_bias = reinterpret_cast<float*>(ptrs_host[4 * j + 3]);
    J.FA = dims_host[5 * j]; J.M_out = dims_host[5 * j + 1]; J.FB = dims_host[5 * j + 2]; J.N_in = dims_host[5 * j + 3];
    J.ldc = dims_host[5 * j + 4];
    VX_REQUIRE(J.A && J.B && J.C, "vx_mlp_dw_batch", "null pointer");
    VX_REQUIRE(J.FA % 32 == 0 && J.FA >= 32 && J.FA <= MLP_MAXW && J.FB % 32 == 0 && J.FB >= 32 && J.FB <= MLP_MAXW &&
               J.M_out >= 1 && J.M_out <= J.FA && J.N_in >= 1 && J.N_in <= J.FB, "vx_mlp_dw_batch", "shape");
    // per-slice cost: MMA columns of both M tiles, or the transposition when that is longer
    const int m_tiles = (J.M_out + MLP_ROWS - 1) / MLP_ROWS;
    cost[j] = m_tiles * (3.0 * J.FB + 32) + 0.3 * (J.FA + J.FB) + 60;
    total += cost[j];
  }
  // deal the SMs to the jobs in proportion to their cost (largest remainder), at most one CTA per 4 slices
  const int sms = vx_num_sms();
  const int cap_per_job = max(1, (slices_cap + 3) / 4);
  int given = 0, n_cta[DW_MAX_JOBS];
  for (int j = 0; j < n_jobs; ++j) {
    n_cta[j] = max(1, (int)(sms * cost[j] / total));
    given += n_cta[j];
  }
  for (int j = 0; given < sms; j = (j + 1) % n_jobs) { ++n_cta[j]; ++given; }   // (given may exceed sms by < n_jobs: harmless)
  int begin = 0;
  for (int j = 0; j < n_jobs; ++j) {
    b.job[j].cta_begin = begin;
    b.job[j].cta_count = min(n_cta[j], cap_per_job);
    begin += b.job[j].cta_count;
  }
  static bool attr_set = false;
  const int smem = (int)sizeof(DwmSmem) + 1024;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { vx_set_error("vx_mlp_dw_batch", cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  k_mlp_dw<<<begin, DW_THREADS, smem, st>>>(b, n_rows_dev, capacity);
  return vx_check_launch("vx_mlp_dw_batch");
}

VX_API int vx_mlp_dw(const float* A_img, int FA, int M_out, const float* B_img, int FB, int N_in, const int* n_rows_dev,
                     int capacity, float* C, int ldc, float* c_bias, cudaStream_t st) {
  const int64_t ptrs[4] = {(int64_t)(uintptr_t)A_img, (int64_t)(uintptr_t)B_img, (int64_t)(uintptr_t)C, (int64_t)(uintptr_t)c_bias};
  const int dims[5] = {FA, M_out, FB, N_in, ldc};
  return vx_mlp_dw_batch(1, ptrs, dims, n_rows_dev, capacity, st);
}

// ---------------------------------------------------------------------------------------------
// Layout probe (tests / development): one M = 128, K = 8 TF32 MMA on caller-provided shared-memory images and
// descriptor fields, D returned as (128, N).  desc_*_fields = the 64-bit smem descriptor without its start address.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
k_umma_probe(const float* __restrict__ A_img, int a_floats, const float* __restrict__ B_img, int b_floats,
             uint64_t desc_a, uint64_t desc_b, uint32_t idesc, int N, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  float* sA = reinterpret_cast<float*>(smem_raw);
  float* sB = reinterpret_cast<float*>(smem_raw + 32768);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < a_floats; i += 128) sA[i] = A_img[i];
  for (int i = tid; i < b_floats; i += 128) sB[i] = B_img[i];
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint64_t da = desc_a | (uint64_t)((smem_u32(sA) >> 4) & 0x3FFF);
    const uint64_t db = desc_b | (uint64_t)((smem_u32(sB) >> 4) & 0x3FFF);
    umma_tf32_ss(tmem, da, db, idesc, 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(lane_addr + c0, v);
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < N) D[tid * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

VX_API int vx_umma_probe(const float* A_img, int a_floats, const float* B_img, int b_floats, int64_t desc_a_fields,
                         int64_t desc_b_fields, int64_t idesc, int N, float* D, cudaStream_t st) {
  VX_REQUIRE(a_floats <= 8192 && b_floats <= 8192 && N % 16 == 0 && N >= 16 && N <= 256, "vx_umma_probe", "sizes");
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
    attr_set = true;
  }
  k_umma_probe<<<1, 128, 65536 + 1024, st>>>(A_img, a_floats, B_img, b_floats, (uint64_t)desc_a_fields, (uint64_t)desc_b_fields,
                                            (uint32_t)idesc, N, D);
  return vx_check_launch("vx_umma_probe");
}

// MMA issue-rate probe (development): `n_mma` back-to-back M = 128, K = 8 TF32 MMAs of width N on zeroed operands, A from
// shared memory (form 0) or from TMEM (form 1), accumulating into one D tile (n_acc = 1) or alternating between two;
// cycles[block] = clock64 ticks from the first issue to the completion of the last MMA.
__global__ void __launch_bounds__(128, 1)
k_umma_rate(int n_mma, int M, int N, int form, int n_acc, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  float* sA = reinterpret_cast<float*>(smem_raw);
  float* sB = reinterpret_cast<float*>(smem_raw + 16384);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 4096; i += 128) sA[i] = 0.f;
  for (int i = tid; i < 8192; i += 128) sB[i] = 0.f;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_tf32(M, N);
    const uint64_t da = make_desc(smem_u32(sA), 128 * 16, 128);
    const uint64_t db = make_desc(smem_u32(sB), N * 16, 128);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t d = tmem + ((n_acc > 1 && (i & 1)) ? 256 : 0);
      if (form == 0) umma_tf32_ss(d, da, db, idesc, 1);
      else umma_tf32_ts(d, tmem + 496, db, idesc, 1);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

VX_API int vx_umma_rate(int n_blocks, int n_mma, int M, int N, int form, int n_acc, int64_t* cycles, cudaStream_t st) {
  VX_REQUIRE((M == 64 || M == 128) && N % 16 == 0 && N >= 16 && N <= 240 && n_blocks >= 1, "vx_umma_rate", "sizes");
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_umma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    attr_set = true;
  }
  k_umma_rate<<<n_blocks, 128, 65536, st>>>(n_mma, M, N, form, n_acc, reinterpret_cast<long long*>(cycles));
  return vx_check_launch("vx_umma_rate");
}
