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
#define MLP_TMEM_COLS 512     // chain kernel: accumulators at columns 0 / 128 (alternating), A (hi) operand at 320

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
//
// One persistent CTA per SM walks a queue of (job, 128-row tile) items; a job is a layer chain of one network (the
// forward chain of rgbnet, the forward chain of k_rgbnet, or one of the two dX chains), up to MLP_MAX_JOBS jobs per
// launch so that the tiles of both networks fill the 148 SMs in one go (706 tile items = 4.8 waves instead of 2 x 3).
//
// Warp roles (576 threads):
//   warps 0-15  row warps.  TMEM lane quarter = warp % 4 (a warp may only touch lanes 32 * (warp % 4) ..), column group
//               cg = warp / 4: thread (row, cg) owns the 16 columns [64 c + 16 cg, +16) of every 64-column chunk c.
//   warp 16     weight producer: lane 0 streams the K = 32 weight slices (hi and lo images) with bulk copies through a
//               2-slot ring, in exactly the order the MMA warp consumes them.
//   warp 17     MMA issuer: lane 0 issues tcgen05.mma (TF32 x 3) and commits.
//
// Pipelining inside a tile.  TMEM (512 columns): A operand (hi part, K <= 192 columns) at 320; the accumulator of MMA
// layer q (a per-CTA running count) at column 0 for even q and 128 for odd q.  Consecutive accumulators overlap only in
// columns [128, 192).  The epilogue of layer q walks its three 64-column chunks starting with the one inside that
// overlap; as soon as a chunk is done -- accumulator columns read, bias / ReLU (or the ReLU gate of the dX chain)
// applied, the hi part written to TMEM A and the lo part to shared memory, the raw values to the HBM row image -- the
// row warps arrive on that chunk's mbarrier and the MMA warp issues the 8 K-steps of layer q + 1 that consume it, into
// the OTHER accumulator, while the row warps work on the next chunk.  So a layer costs ~ (epilogue / 3 + MMAs) instead
// of (epilogue + MMAs).  The same hand-over happens across tiles: in the epilogue of a tile's last MMA layer, right
// after the first chunk, the row warps stage the input rows of the CTA's next tile (prefetched into registers before
// the accumulator wait) into the A operand, so the next tile's first layer runs under the rest of that epilogue.
//
// Forward chains end in a tiny N = 3 layer: 72 MMAs at the 96-clock floor of tcgen05.mma (as much tensor time as a
// 192-wide layer) for 0.2 % of the flops.  It runs on the CUDA cores instead, inside the epilogue of the last hidden
// layer: every thread accumulates its 48 columns' share of the three dot products in exact fp32 (weights broadcast from
// shared memory), the four column groups meet through shared memory.
//
// k_rgbnet's input carries rgbnet's output for the same row (rgb_logit.detach(), lib/voxurf_fine.py:741-751): job 1
// declares `dep = 0` and patches those input columns from job 0's output.  Job 0's tiles come first in the queue and set
// a per-tile flag (release) when their output rows are in memory; the consumer tile spins on it (acquire).  All CTAs are
// co-resident (grid <= #SMs, one CTA per SM) and every CTA handles its job-0 items before its job-1 items: no deadlock.
// ---------------------------------------------------------------------------------------------
#define MC_ROW_WARPS 16
#define MC_ROW_THREADS (MC_ROW_WARPS * 32)
#define MC_THREADS (MC_ROW_THREADS + 96)     // + producer warp, MMA warp, publisher warp
#define MC_TMEM_A 320
#define MLP_MAX_JOBS 2
#define MC_MAX_FINAL 4          // outputs of the CUDA-core final layer

struct MlpLayer {
  const float* W_hi;    // CH(Np) image, Kp/4 chunks
  const float* W_lo;
  const float* bias;    // [N] or nullptr
  float* img;           // ACT(Np) row image of this layer's output, or nullptr
  const unsigned long long* gate;  // ReLU gates applied to this layer's output (dX chains), one 64-bit word per (row, column
                        //   group cg): feature f = 64 c + 16 cg + j <-> bit 16 c + j of gate[r * 4 + cg] (1 = pass); nullptr = none
  unsigned long long* gate_out;    // the same words written for this layer's own output (forward chains: output > 0), or nullptr
  int Kp, Np, N, relu;
};
struct MlpJob {
  const float* X;       // (capacity, ldx) input rows, K0 valid columns
  float* x_img;         // ACT(pad32(K0p)) row image of the input, or nullptr
  float* Y;             // (capacity, ldy) output rows, n_out columns
  const float* Wf;      // final layer on the CUDA cores: Y[r][o] = bias_f[o] + sum_k act[r][k] * Wf[o * ldwf + k]; nullptr = none
  const float* bias_f;  //   (then the last MMA layer's output is Y itself)
  const float* patch;   // input columns [patch_col, patch_col + patch_n) are read from patch[r * patch_ld + j] instead of X
  int ldx, K0, K0p, ldy, n_out, ldwf, patch_col, patch_n, patch_ld, dep, n_layers;
  MlpLayer L[MLP_MAX_LAYERS];
};
struct MlpBatch {
  int n_jobs;
  int* done;            // per (job with dependents, tile) completion flags, zeroed by the launcher; may be nullptr
  long long* trace;     // development (-DMC_TRACE builds): CTA 0 logs (clock64 << 8 | code) of its pipeline events, else nullptr
  MlpJob job[MLP_MAX_JOBS];
};
#ifdef MC_TRACE
#define MC_T(base, idx, code) do { if (batch.trace && blockIdx.x == 0 && lane == 0 && (idx) < 2040) batch.trace[(base) + (idx)++] = (clock64() << 8) | (code); } while (0)
#else
#define MC_T(base, idx, code) do { } while (0)
#endif
static long long* g_mc_trace = nullptr;

struct __align__(16) MlpSmem {
  float A_lo[MLP_MAXW / 4 * MLP_ROWS * 4];                           // 96 KB  [K/4][128][4]
  float B[MLP_STAGES][2][MLP_SLICE_K / 4 * MLP_MAXW * 4];            // stages x {hi,lo} x [8 chunks][192][4] = 2 x 48 KB
  float bias[MLP_MAX_JOBS][MLP_MAX_LAYERS][MLP_MAXW];
  float Wf[MLP_MAX_JOBS][MC_MAX_FINAL][MLP_MAXW];
  float bias_f[MLP_MAX_JOBS][MC_MAX_FINAL];
  float red[4][MLP_ROWS][MC_MAX_FINAL];                               // final-layer partial sums of the 4 column groups
  uint64_t bar_full[MLP_STAGES];
  uint64_t bar_empty[MLP_STAGES];
  uint64_t bar_chunk[3];
  uint64_t bar_acc;
  uint64_t bar_pub;
  uint32_t tmem_base;
};

__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_rows16() { asm volatile("bar.sync 1, 512;" ::: "memory"); }   // the 512 row threads only

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

// TMEM load split into issue and wait so that the next chunk's accumulator columns fly while this chunk is processed;
// the wait names the destination registers as in/out operands, which orders every later use of them behind it
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16_wait(float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}

// chunk visiting order of an epilogue whose accumulator sits at column 0 (even) / 128 (odd): the chunk shared with the
// next accumulator first
__device__ __forceinline__ int mc_chunk(int odd, int i) { return odd ? i : (i == 0 ? 2 : i - 1); }

// 16 consecutive features [c0, c0 + 16) of this thread's row become A-operand columns: hi -> TMEM, lo -> smem tile; the
// raw values optionally go to the HBM row image (two aligned 32-byte pieces = one contiguous 64-byte run)
__device__ __forceinline__ void mc_store_a16(MlpSmem& s, uint32_t tmem_a_lane, int rt, int c0, const float* v,
                                             float* __restrict__ img, int F, int64_t row) {
  // The A operand's hi part is x itself: tcgen05.mma.kind::tf32 reads the upper 19 bits of each 32-bit element, i.e. x
  // truncated to TF32, and lo = x - trunc(x) is exact in fp32 (it is truncated to TF32 in turn: |error| <= 2^-21 |x|, the
  // same order as the lo * lo term the 3-product split drops; tests/test_gpu_mlp.py holds the 5e-6 / 2e-5 bars).  Two
  // instructions per element instead of five for the round-to-nearest split.
#ifndef EXP_NO_TMEMST
  tmem_st16(tmem_a_lane + c0, v);
#endif
  float lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) lo[j] = v[j] - __uint_as_float(__float_as_uint(v[j]) & 0xffffe000u);
#ifndef EXP_NO_LO
#pragma unroll
  for (int q = 0; q < 4; ++q)
    reinterpret_cast<float4*>(s.A_lo)[((c0 >> 2) + q) * MLP_ROWS + rt] = make_float4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
#endif
#ifdef EXP_NO_IMG
  img = nullptr;
#endif
  if (img) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(img + act_offset(row, c0 + 8 * h, F)),
                   "f"(v[8 * h]), "f"(v[8 * h + 1]), "f"(v[8 * h + 2]), "f"(v[8 * h + 3]), "f"(v[8 * h + 4]), "f"(v[8 * h + 5]),
                   "f"(v[8 * h + 6]), "f"(v[8 * h + 7]) : "memory");
  }
}

// this thread's pieces of an input row: piece u covers columns [16 cg + 64 u, +16), u = 0, 1 (K0p <= 128).  Piece 0 is
// prefetched into registers one layer ahead; piece 1 (only the 79-wide rgbnet input has one, for cg = 0) is loaded when
// it is staged -- holding it too costs 16 more live registers in the busiest loop, and 576 threads leave 96 per thread.
struct McRowRegs { float v[16]; };

__device__ __forceinline__ void mc_load_piece(const MlpJob& J, int64_t row, int n_rows, int c0, float* v, bool patch_ready) {
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = 0.f;
  if (c0 >= J.K0p || row >= n_rows) return;
  const float* src = J.X + row * J.ldx;
  if ((J.ldx & 3) == 0 && c0 + 16 <= J.K0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src + c0 + 4 * q));
      v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < J.K0) v[j] = __ldg(src + c0 + j);
  }
  (void)patch_ready;
}

// columns produced by another job of this launch (after its completion flag was acquired): plain L2 loads, never the
// read-only path
__device__ __forceinline__ void mc_patch_piece(const MlpJob& J, int64_t row, int n_rows, int c0, float* v) {
  if (!J.patch || row >= n_rows || c0 >= J.K0p) return;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int p = c0 + j - J.patch_col;
    if (p >= 0 && p < J.patch_n) {
      float x;
      asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(x) : "l"(J.patch + row * J.patch_ld + p));
      v[j] = x;
    }
  }
}

__global__ void __launch_bounds__(MC_THREADS, 1)
k_mlp_chain(const __grid_constant__ MlpBatch batch, const int* __restrict__ n_rows_dev, int capacity) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  MlpSmem& s = *reinterpret_cast<MlpSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_rows = min(*n_rows_dev, capacity);
  const int n_tiles = (n_rows + MLP_ROWS - 1) / MLP_ROWS;
  const int n_items = n_tiles * batch.n_jobs;      // item u: job u / n_tiles, tile u % n_tiles

  if (tid == 0) {
    for (int i = 0; i < MLP_STAGES; ++i) { mbar_init(&s.bar_full[i], 1); mbar_init(&s.bar_empty[i], 1); }
    for (int i = 0; i < 3; ++i) mbar_init(&s.bar_chunk[i], MC_ROW_WARPS);
    mbar_init(&s.bar_acc, 1);
    mbar_init(&s.bar_pub, 4);
    fence_barrier_init();
  }
  for (int j = 0; j < batch.n_jobs; ++j) {
    const MlpJob& J = batch.job[j];
    for (int l = 0; l < J.n_layers; ++l)
      for (int c = tid; c < MLP_MAXW; c += MC_THREADS) s.bias[j][l][c] = (J.L[l].bias && c < J.L[l].N) ? J.L[l].bias[c] : 0.f;
    if (J.Wf) {
      const int Kf = J.L[J.n_layers - 1].N;
      for (int i = tid; i < MC_MAX_FINAL * MLP_MAXW; i += MC_THREADS) {
        const int o = i / MLP_MAXW, k = i % MLP_MAXW;
        s.Wf[j][o][k] = (o < J.n_out && k < Kf) ? J.Wf[(int64_t)o * J.ldwf + k] : 0.f;
      }
      if (tid < MC_MAX_FINAL) s.bias_f[j][tid] = (J.bias_f && tid < J.n_out) ? J.bias_f[tid] : 0.f;
    }
  }
  if (warp == 0) tmem_alloc(&s.tmem_base, MLP_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s.tmem_base;

  if (warp == MC_ROW_WARPS) {
    // ===== weight producer: every (item, layer, slice) weight block in consumption order, two slices ahead at most =====
    if (lane == 0) {
      uint32_t it = 0, q = 0;
      for (int u = blockIdx.x; u < n_items; u += gridDim.x) {
        const MlpJob& J = batch.job[u / n_tiles];
        for (int l = 0; l < J.n_layers; ++l, ++q) {
          const MlpLayer& L = J.L[l];
          const int NSL = (L.Kp + MLP_SLICE_K - 1) / MLP_SLICE_K;
          const int prev_odd = (int)((q - 1) & 1);     // order of the chunks of the A operand = order the previous epilogue made them
          for (int i = 0; i < 3; ++i) {
            const int c = (l == 0) ? i : mc_chunk(prev_odd, i);
            for (int sl = 2 * c; sl < 2 * c + 2 && sl < NSL; ++sl, ++it) {
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
      }
    }
  } else if (warp == MC_ROW_WARPS + 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t it = 0, q = 0;
      int ti = 0; (void)ti;
      for (int u = blockIdx.x; u < n_items; u += gridDim.x) {
        const MlpJob& J = batch.job[u / n_tiles];
        for (int l = 0; l < J.n_layers; ++l, ++q) {
          const MlpLayer& L = J.L[l];
          const int KS = L.Kp / 8, Np = L.Np;
          const int NSL = (L.Kp + MLP_SLICE_K - 1) / MLP_SLICE_K;
          const uint32_t d_tm = tmem + ((q & 1) ? 128u : 0u);
          const uint32_t idesc = make_idesc_tf32(MLP_ROWS, Np);
          const int prev_odd = (int)((q - 1) & 1);
          uint32_t first = 1;
          for (int i = 0; i < 3; ++i) {
            const int c = (l == 0) ? i : mc_chunk(prev_odd, i);
            mbar_wait(&s.bar_chunk[c], q & 1);        // A columns [64 c, 64 c + 64) are in place (and the accumulator
            tc_fence_after();                         // columns this layer shares with the previous one have been read)
            MC_T(2048, ti, 40 + i);
            for (int sl = 2 * c; sl < 2 * c + 2 && sl < NSL; ++sl, ++it) {
              const int slot = it % MLP_STAGES;
              MC_T(2048, ti, 60);
              mbar_wait(&s.bar_full[slot], (it / MLP_STAGES) & 1);
              tc_fence_after();
              MC_T(2048, ti, 61);
              const int k_steps = min(MLP_SLICE_K / 8, KS - sl * (MLP_SLICE_K / 8));
              for (int kk = 0; kk < k_steps; ++kk) {
                const int ks = sl * (MLP_SLICE_K / 8) + kk;
                const uint32_t a_tm = tmem + MC_TMEM_A + ks * 8;
                const uint64_t da_lo = make_desc(smem_u32(s.A_lo) + (uint32_t)(ks * 2) * MLP_ROWS * 16, MLP_ROWS * 16, 128);
                const uint32_t b_off = (uint32_t)(kk * 2) * Np * 16;
                const uint64_t db_hi = make_desc(smem_u32(&s.B[slot][0][0]) + b_off, Np * 16, 128);
                const uint64_t db_lo = make_desc(smem_u32(&s.B[slot][1][0]) + b_off, Np * 16, 128);
                umma_tf32_ts(d_tm, a_tm, db_hi, idesc, first ^ 1u);
                umma_tf32_ts(d_tm, a_tm, db_lo, idesc, 1);
                umma_tf32_ss(d_tm, da_lo, db_hi, idesc, 1);
                first = 0;
              }
              umma_commit(&s.bar_empty[slot]);
            }
          }
          umma_commit(&s.bar_acc);
          MC_T(2048, ti, 50);
        }
      }
    }
  } else if (warp == MC_ROW_WARPS + 2) {
    // ===== publisher: releases the completion flag of every tile whose output rows other jobs read.  The four warps
    // that write a tile's output rows arrive on bar_pub (release.cta); the gpu-scope release store then happens here,
    // so that no row warp sits behind a memory fence =====
    if (lane == 0 && batch.done) {
      uint32_t n_pub = 0;
      for (int u = blockIdx.x; u < n_items; u += gridDim.x) {
        const int ji = u / n_tiles, tile = u % n_tiles;
        if (ji + 1 >= batch.n_jobs) continue;
        mbar_wait(&s.bar_pub, n_pub & 1);
        ++n_pub;
        int one = 1;
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(batch.done + ji * n_tiles + tile), "r"(one) : "memory");
      }
    }
  } else {
    // ===== row warps =====
    const int rt = (warp & 3) * 32 + lane;       // row within the tile == TMEM lane
    const int cg = warp >> 2;                    // column group
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t a_lane = lane_addr + MC_TMEM_A;
    uint32_t q = 0;
    int ti = 0; (void)ti;

    // stage the input rows of item u as the A operand of its first layer and release the three chunk barriers
    auto stage = [&](int u, McRowRegs& R, bool loaded) {
      const int ji = u / n_tiles, tile = u % n_tiles;
      const MlpJob& J = batch.job[ji];
      const int64_t row = (int64_t)tile * MLP_ROWS + rt;
      const int F = (J.K0p + 31) & ~31;
      // every load first (piece 0 unless it was prefetched, piece 1 if the row is wider than 64, the patched columns), then
      // the stores: one memory latency instead of three
      const bool two = 16 * cg + 64 < J.K0p;          // warp-uniform
      float w[16];
      if (!loaded) mc_load_piece(J, row, n_rows, 16 * cg, R.v, false);
      if (two) mc_load_piece(J, row, n_rows, 16 * cg + 64, w, false);
      if (J.patch) {
        if (J.dep >= 0 && batch.done) {
          if (tid == 0) {
            const int* flag = batch.done + J.dep * n_tiles + tile;
            int f;
            do { asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(flag) : "memory"); } while (f == 0);
          }
          bar_rows16();
        }
        mc_patch_piece(J, row, n_rows, 16 * cg, R.v);
        if (two) mc_patch_piece(J, row, n_rows, 16 * cg + 64, w);
      }
      if (16 * cg < J.K0p) mc_store_a16(s, a_lane, rt, 16 * cg, R.v, J.x_img, F, row);
      if (two) mc_store_a16(s, a_lane, rt, 16 * cg + 64, w, J.x_img, F, row);
      tmem_st_wait();
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive_cta(&s.bar_chunk[0]); mbar_arrive_cta(&s.bar_chunk[1]); mbar_arrive_cta(&s.bar_chunk[2]); }
      if (warp == 0) MC_T(0, ti, 20);
    };

    McRowRegs R;
    if ((int)blockIdx.x < n_items) stage(blockIdx.x, R, false);
    for (int u = blockIdx.x; u < n_items; u += gridDim.x) {
      const int ji = u / n_tiles, tile = u % n_tiles;
      const MlpJob& J = batch.job[ji];
      const int64_t row = (int64_t)tile * MLP_ROWS + rt;
      const bool valid = row < n_rows;
      const int u_next = u + gridDim.x;
      for (int l = 0; l < J.n_layers; ++l, ++q) {
        const MlpLayer& L = J.L[l];
        const int Np = L.Np;
        const bool last = (l == J.n_layers - 1);
        const int odd = (int)(q & 1);
        const uint32_t d_lane = lane_addr + (odd ? 128u : 0u);
        if (last && u_next < n_items) {
          // prefetch the next item's input rows: the loads fly while this layer's MMAs finish
          const MlpJob& Jn = batch.job[u_next / n_tiles];
          const int64_t rown = (int64_t)(u_next % n_tiles) * MLP_ROWS + rt;
          mc_load_piece(Jn, rown, n_rows, 16 * cg, R.v, false);
        }
        // ReLU gates of the dX chain: 16 bits per chunk visit, fetched while the MMAs run (the forward chain left them as a
        // 192-bit bitmap per row: 24 bytes instead of the 768-byte activation row)
        unsigned long long gbits = ~0ull, gout = 0;    // 16 bits per chunk c
        if (L.gate && valid) gbits = __ldg(L.gate + row * 4 + cg);
        if (warp == 0) MC_T(0, ti, 2);
        mbar_wait(&s.bar_acc, q & 1);
        tc_fence_after();
        if (warp == 0) MC_T(0, ti, 1);
        float acc[MC_MAX_FINAL];
#pragma unroll
        for (int o = 0; o < MC_MAX_FINAL; ++o) acc[o] = 0.f;
        // per-layer constants out of the (dynamically indexed) kernel-parameter structs
        float* const img = L.img;
        unsigned long long* const gate_out = L.gate_out;
        const bool has_gate = L.gate != nullptr, relu = L.relu != 0, has_bias = L.bias != nullptr, has_wf = J.Wf != nullptr;
        const float validf = valid ? 1.f : 0.f;      // rows past n_rows: zero A rows give zero accumulators; only the bias must go
        const float* const bias_s = s.bias[ji][l];
        const int c00 = 64 * mc_chunk(odd, 0) + 16 * cg, c01 = 64 * mc_chunk(odd, 1) + 16 * cg, c02 = 64 * mc_chunk(odd, 2) + 16 * cg;
        // Last layer of the item: the next item's first layer may start as soon as the accumulator columns it shares with
        // this one ([128, 192): the first chunk in visiting order) have been READ -- so read them, stage the next item's
        // input rows as the A operand (the MMA warp takes over from there), and only then do the arithmetic on them.
        float vn[16];
        const bool early = last && u_next < n_items;
        bool pend = false;                           // a TMEM load into vn is in flight
        if (c00 < Np) { tmem_ld16_issue(d_lane + c00, vn); pend = true; }
        if (early) {
          if (c00 < Np) tmem_ld16_wait(vn);
          tc_fence_before();
          stage(u_next, R, true);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int c0 = (i == 0) ? c00 : (i == 1 ? c01 : c02);
          if (c0 < Np) {                       // warp-uniform
            float v[16];
            if (warp == 0) MC_T(0, ti, 70);
            if (!pend) tmem_ld16_issue(d_lane + c0, vn);
            tmem_ld16_wait(vn);
            pend = false;
            if (warp == 0) MC_T(0, ti, 71);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = vn[j];
            if (i < 2) {                       // next visit's accumulator columns fly while this visit is processed
              const int c1 = (i == 0) ? c01 : c02;
              if (c1 < Np) { tmem_ld16_issue(d_lane + c1, vn); pend = true; }
            }
#ifdef EXP_NO_BIAS
            if (false) {
#else
            if (has_bias) {
#endif
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * q4);
                v[4 * q4] = fmaf(b4.x, validf, v[4 * q4]); v[4 * q4 + 1] = fmaf(b4.y, validf, v[4 * q4 + 1]);
                v[4 * q4 + 2] = fmaf(b4.z, validf, v[4 * q4 + 2]); v[4 * q4 + 3] = fmaf(b4.w, validf, v[4 * q4 + 3]);
              }
            }
            if (relu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (has_gate) {
              const uint32_t gb = (uint32_t)(gbits >> (c0 >> 2 & 0x30));     // chunk c = c0 / 64 -> bits 16 c ..
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = (gb & (1u << j)) ? v[j] : 0.f;
            }
#ifdef EXP_NO_GATE
            if (false) {
#else
            if (gate_out) {
#endif
              // bit j = (v[j] > 0); v >= 0 here (post-ReLU), so v > 0 <=> its bit pattern is non-zero: adding 0x7fffffff carries
              // into bit 31 exactly then, and a funnel shift collects that bit
              uint32_t gb = 0;
#pragma unroll
              for (int j = 15; j >= 0; --j) gb = __funnelshift_l(__float_as_uint(v[j]) + 0x7fffffffu, gb, 1);
              gout |= (unsigned long long)gb << (c0 >> 2 & 0x30);
            }
            if (warp == 0) MC_T(0, ti, 72);
            if (!last) {
              mc_store_a16(s, a_lane, rt, c0, v, img, Np, row);
              if (warp == 0) MC_T(0, ti, 73);
            } else {
              if (img) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
                  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(img + act_offset(row, c0 + 8 * h, Np)),
                               "f"(v[8 * h]), "f"(v[8 * h + 1]), "f"(v[8 * h + 2]), "f"(v[8 * h + 3]), "f"(v[8 * h + 4]), "f"(v[8 * h + 5]),
                               "f"(v[8 * h + 6]), "f"(v[8 * h + 7]) : "memory");
              }
              if (has_wf) {
#pragma unroll
                for (int o = 0; o < MC_MAX_FINAL; ++o) {
#pragma unroll
                  for (int q4 = 0; q4 < 4; ++q4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(&s.Wf[ji][o][c0 + 4 * q4]);
                    acc[o] = fmaf(v[4 * q4], w4.x, acc[o]); acc[o] = fmaf(v[4 * q4 + 1], w4.y, acc[o]);
                    acc[o] = fmaf(v[4 * q4 + 2], w4.z, acc[o]); acc[o] = fmaf(v[4 * q4 + 3], w4.w, acc[o]);
                  }
                }
              } else if (valid) {
                float* dst = J.Y + row * J.ldy + c0;
                if ((J.ldy & 3) == 0 && c0 + 16 <= J.n_out) {
#pragma unroll
                  for (int qv = 0; qv < 4; ++qv)
                    reinterpret_cast<float4*>(dst)[qv] = make_float4(v[4 * qv], v[4 * qv + 1], v[4 * qv + 2], v[4 * qv + 3]);
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (c0 + j < J.n_out) dst[j] = v[j];
                }
              }
            }
          }
          if (!last) {
            // this chunk of the next layer's A operand is complete (and this chunk of the accumulator has been read)
#ifndef EXP_NO_FENCE
            tmem_st_wait();
            if (warp == 0) MC_T(0, ti, 74);
            fence_proxy_async();
#endif
            tc_fence_before();
            if (warp == 0) MC_T(0, ti, 75);
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(&s.bar_chunk[(i == 0) ? mc_chunk(odd, 0) : (i == 1 ? mc_chunk(odd, 1) : mc_chunk(odd, 2))]);
            if (warp == 0) MC_T(0, ti, 10 + i);
          }
        }
#ifndef EXP_NO_GSTORE
        if (gate_out) gate_out[row * 4 + cg] = gout;      // one 8-byte store per thread and layer
#endif
        if (last && J.Wf) {
          // CUDA-core final layer: the four column groups of a row meet in shared memory
#pragma unroll
          for (int o = 0; o < MC_MAX_FINAL; ++o) s.red[cg][rt][o] = acc[o];
          bar_rows16();
          if (cg == 0) {
            if (valid) {
              for (int o = 0; o < J.n_out; ++o)
                J.Y[row * J.ldy + o] = ((s.red[0][rt][o] + s.red[1][rt][o]) + (s.red[2][rt][o] + s.red[3][rt][o])) + s.bias_f[ji][o];
            }
            if (batch.done && ji + 1 < batch.n_jobs) {   // this tile's output rows are written: hand the flag to the publisher
              __syncwarp();
              if (lane == 0) mbar_arrive_cta(&s.bar_pub);
            }
          }
          bar_rows16();     // s.red is reused by the next item
        }
        if (last && warp == 0) MC_T(0, ti, 30);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, MLP_TMEM_COLS);
}

// Jobs are described by packed HOST arrays so the C ABI stays plain.  Per job j:
//   ptrs_host[j*30 + {0..5}]  = X, x_img, Y, Wf, bias_f, patch                       (device addresses, 0 = none)
//   ptrs_host[j*30 + 6 + l*6 + {0..5}] = layer l: W_hi, W_lo (CH(Np) images from vx_mlp_prep), bias, row image out,
//                                        gate bitmap in, gate bitmap out
//   dims_host[j*26 + {0..9}]  = ldx, K0, n_layers, ldy, n_out, ldwf, patch_col, patch_n, patch_ld, dep (-1 = none)
//   dims_host[j*26 + 10 + l*4 + {0..3}] = layer l: Kp, Np, N, relu
// done_flags: device int array of n_jobs * ceil(capacity / 128) entries (needed when a job has dep >= 0), zeroed here.
// Every row image needs 128 * ceil(capacity / 128) rows.
#define MC_JOB_STRIDE 26      // dims per job
#define MC_PTR_STRIDE 30      // pointers per job
VX_API int vx_mlp_chain_batch(int n_jobs, const int64_t* ptrs_host, const int* dims_host, const int* n_rows_dev, int capacity,
                              int* done_flags, cudaStream_t st) {
  VX_REQUIRE(n_jobs >= 1 && n_jobs <= MLP_MAX_JOBS, "vx_mlp_chain_batch", "1..2 jobs");
  VX_REQUIRE(n_rows_dev != nullptr, "vx_mlp_chain_batch", "n_rows_dev required");
  MlpBatch b;
  memset(&b, 0, sizeof(b));
  b.n_jobs = n_jobs;
  b.trace = g_mc_trace;
  bool any_dep = false;
  for (int j = 0; j < n_jobs; ++j) {
    MlpJob& J = b.job[j];
    const int64_t* P = ptrs_host + j * MC_PTR_STRIDE;
    const int* D = dims_host + j * MC_JOB_STRIDE;
    J.X = reinterpret_cast<const float*>(P[0]);
    J.x_img = reinterpret_cast<float*>(P[1]);
    J.Y = reinterpret_cast<float*>(P[2]);
    J.Wf = reinterpret_cast<const float*>(P[3]);
    J.bias_f = reinterpret_cast<const float*>(P[4]);
    J.patch = reinterpret_cast<const float*>(P[5]);
    J.ldx = D[0]; J.K0 = D[1]; J.n_layers = D[2]; J.ldy = D[3]; J.n_out = D[4]; J.ldwf = D[5];
    J.patch_col = D[6]; J.patch_n = D[7]; J.patch_ld = D[8]; J.dep = D[9];
    VX_REQUIRE(J.n_layers >= 1 && J.n_layers <= MLP_MAX_LAYERS, "vx_mlp_chain_batch", "1..4 MMA layers");
    VX_REQUIRE(J.X && J.Y, "vx_mlp_chain_batch", "X / Y required");
    for (int l = 0; l < J.n_layers; ++l) {
      MlpLayer& L = J.L[l];
      L.W_hi = reinterpret_cast<const float*>(P[6 + l * 6 + 0]);
      L.W_lo = reinterpret_cast<const float*>(P[6 + l * 6 + 1]);
      L.bias = reinterpret_cast<const float*>(P[6 + l * 6 + 2]);
      L.img = reinterpret_cast<float*>(P[6 + l * 6 + 3]);
      L.gate = reinterpret_cast<const unsigned long long*>(P[6 + l * 6 + 4]);
      L.gate_out = reinterpret_cast<unsigned long long*>(P[6 + l * 6 + 5]);
      L.Kp = D[10 + l * 4 + 0]; L.Np = D[10 + l * 4 + 1]; L.N = D[10 + l * 4 + 2]; L.relu = D[10 + l * 4 + 3];
      VX_REQUIRE(L.W_hi && L.W_lo, "vx_mlp_chain_batch", "weight images required");
      VX_REQUIRE(L.Kp % 8 == 0 && L.Kp >= 8 && L.Kp <= MLP_MAXW && L.Np % 16 == 0 && L.Np >= 16 && L.Np <= MLP_MAXW,
                 "vx_mlp_chain_batch", "layer shape");
      if (l + 1 < J.n_layers)
        VX_REQUIRE(L.Np % 32 == 0 && L.Np == D[10 + (l + 1) * 4 + 0], "vx_mlp_chain_batch", "hidden widths must chain and be multiples of 32");
      VX_REQUIRE(!L.img || L.Np % 32 == 0, "vx_mlp_chain_batch", "a row image needs a width that is a multiple of 32");
      VX_REQUIRE(!(L.gate || L.gate_out) || L.Np % 16 == 0, "vx_mlp_chain_batch", "gate words need a width that is a multiple of 16");
    }
    J.K0p = J.L[0].Kp;
    const MlpLayer& LL = J.L[J.n_layers - 1];
    VX_REQUIRE(J.K0 <= J.K0p && J.K0 <= J.ldx && J.K0p <= 128, "vx_mlp_chain_batch", "K0");
    if (J.Wf) VX_REQUIRE(J.n_out >= 1 && J.n_out <= MC_MAX_FINAL && J.ldwf >= LL.N, "vx_mlp_chain_batch", "final layer: 1..4 outputs");
    else VX_REQUIRE(J.n_out <= LL.Np, "vx_mlp_chain_batch", "n_out");
    if (J.patch) {
      VX_REQUIRE(J.patch_n >= 1 && J.patch_col >= 0 && J.patch_col + J.patch_n <= J.K0 && J.dep < j, "vx_mlp_chain_batch", "patch");
      if (J.dep >= 0) {
        any_dep = true;
        VX_REQUIRE(b.job[J.dep].Wf != nullptr && J.dep + 1 < n_jobs, "vx_mlp_chain_batch", "a job others patch from must end in the CUDA-core final layer");
      }
    }
  }
  const int tiles_cap = (capacity + MLP_ROWS - 1) / MLP_ROWS;
  if (tiles_cap <= 0) return 0;
  if (any_dep) {
    VX_REQUIRE(done_flags != nullptr, "vx_mlp_chain_batch", "done_flags required when a job patches its input from another");
    b.done = done_flags;
    cudaError_t e = cudaMemsetAsync(done_flags, 0, sizeof(int) * (size_t)n_jobs * tiles_cap, st);
    if (e != cudaSuccess) { vx_set_error("vx_mlp_chain_batch", cudaGetErrorString(e)); return (int)e; }
  }
  static bool attr_set = false;
  const int smem = (int)sizeof(MlpSmem) + 1024;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { vx_set_error("vx_mlp_chain_batch", cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  const int blocks = min(tiles_cap * n_jobs, vx_num_sms());
  k_mlp_chain<<<blocks, MC_THREADS, smem, st>>>(b, n_rows_dev, capacity);
  return vx_check_launch("vx_mlp_chain_batch");
}

// development: device buffer of 4096 int64 that CTA 0 of the next chain launches logs its pipeline events into (only in
// -DMC_TRACE builds; nullptr switches it off)
VX_API int vx_mlp_trace_set(int64_t* buf) {
  g_mc_trace = reinterpret_cast<long long*>(buf);
  return 0;
}

// One chain, every layer on the tensor cores (tests, generic callers): layer l described by
//   ptrs_host[l*5 + {0..4}] = W_hi, W_lo, bias, row image out, gate bitmap in;  dims_host[l*4 + {0..3}] = Kp, Np, N, relu
VX_API int vx_mlp_chain(const float* X, int ldx, int K0, const int* n_rows_dev, int capacity, int n_layers,
                        const int64_t* ptrs_host, const int* dims_host, float* Y, int ldy, int n_out, float* x_img,
                        cudaStream_t st) {
  VX_REQUIRE(n_layers >= 1 && n_layers <= MLP_MAX_LAYERS, "vx_mlp_chain", "1..4 layers");
  int64_t P[MC_PTR_STRIDE];
  int D[MC_JOB_STRIDE];
  memset(P, 0, sizeof(P));
  memset(D, 0, sizeof(D));
  P[0] = (int64_t)(uintptr_t)X; P[1] = (int64_t)(uintptr_t)x_img; P[2] = (int64_t)(uintptr_t)Y;
  D[0] = ldx; D[1] = K0; D[2] = n_layers; D[3] = ldy; D[4] = n_out; D[9] = -1;
  for (int l = 0; l < n_layers; ++l) {
    for (int k = 0; k < 5; ++k) P[6 + l * 6 + k] = ptrs_host[l * 5 + k];
    for (int k = 0; k < 4; ++k) D[10 + l * 4 + k] = dims_host[l * 4 + k];
  }
  return vx_mlp_chain_batch(1, P, D, n_rows_dev, capacity, nullptr, st);
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
  // deterministic mode: partial sums go into 64-bit fixed-point accumulators (acc[i] <-> grad_base[i], `scale` units per
  // 1.0) instead of fp32 atomics on C / c_bias: integer sums do not depend on the order the CTAs finish in
  unsigned long long* acc;
  const float* grad_base;
  double scale;
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
          if (m < J.M_out && batch.acc) {
            if (c0 == FB) {
              if (J.c_bias) atomicAdd(batch.acc + (J.c_bias + m - batch.grad_base), (unsigned long long)__double2ll_rn((double)v[0] * batch.scale));
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < J.N_in)
                  atomicAdd(batch.acc + (J.C + (int64_t)m * J.ldc + c0 + j - batch.grad_base), (unsigned long long)__double2ll_rn((double)v[j] * batch.scale));
            }
          } else if (m < J.M_out) {
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
static int dw_batch_impl(int n_jobs, const int64_t* ptrs_host, const int* dims_host, const int* n_rows_dev, int capacity,
                         int64_t* acc, const float* grad_base, float acc_scale, cudaStream_t st);

VX_API int vx_mlp_dw_batch(int n_jobs, const int64_t* ptrs_host, const int* dims_host, const int* n_rows_dev, int capacity,
                           cudaStream_t st) {
  return dw_batch_impl(n_jobs, ptrs_host, dims_host, n_rows_dev, capacity, nullptr, nullptr, 0.f, st);
}

// deterministic variant: every C / c_bias of the jobs lies inside one gradient buffer starting at grad_base; its partial
// sums are added to acc[offset] as 64-bit fixed point (acc_scale units per 1.0); vx_fx_accumulate folds acc into the buffer
VX_API int vx_mlp_dw_batch_fx(int n_jobs, const int64_t* ptrs_host, const int* dims_host, const int* n_rows_dev, int capacity,
                              int64_t* acc, const float* grad_base, float acc_scale, cudaStream_t st) {
  VX_REQUIRE(acc && grad_base && acc_scale > 0.f, "vx_mlp_dw_batch_fx", "acc / grad_base / acc_scale required");
  return dw_batch_impl(n_jobs, ptrs_host, dims_host, n_rows_dev, capacity, acc, grad_base, acc_scale, st);
}

static int dw_batch_impl(int n_jobs, const int64_t* ptrs_host, const int* dims_host, const int* n_rows_dev, int capacity,
                         int64_t* acc, const float* grad_base, float acc_scale, cudaStream_t st) {
  VX_REQUIRE(n_rows_dev != nullptr, "vx_mlp_dw_batch", "n_rows_dev required");
  VX_REQUIRE(n_jobs >= 0 && n_jobs <= DW_MAX_JOBS, "vx_mlp_dw_batch", "at most 8 jobs per launch");
  const int slices_cap = (capacity + DW_KC - 1) / DW_KC;
  if (n_jobs == 0 || slices_cap <= 0) return 0;
  DwBatch b;
  memset(&b, 0, sizeof(b));
  b.n_jobs = n_jobs;
  b.acc = reinterpret_cast<unsigned long long*>(acc);
  b.grad_base = grad_base;
  b.scale = (double)acc_scale;
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
