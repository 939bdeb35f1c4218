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
//   * weights (hi/lo, prepared once per optimizer step by k_mlp_prep) are streamed from L2 in K=32 slices through a
//     double-buffered cp.async ring; one thread issues the MMAs and frees ring slots with tcgen05.commit -> mbarrier;
//   * the hidden activations are also written to HBM when `save_h` (needed by the backward pass).
// The row count is read from device memory (sync-free pipeline); rows past it are computed as zeros.
#include "common.cuh"

#define MLP_ROWS 128
#define MLP_MAXW 192          // max layer width (N and K)
#define MLP_STAGES 2
#define MLP_SLICE_K 32        // K extent of one weight slice in the cp.async ring (4 MMA K-steps)
#define MLP_MAX_LAYERS 4
#define MLP_TMEM_COLS 512     // D accumulator at column 0, A (hi) operand at column 256
#define MLP_TMEM_A 256

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
// weight preparation: hi / lo parts, zero-padded to (Np, Kp); optionally transposed (for the dX chain)
// ---------------------------------------------------------------------------------------------
__global__ void k_mlp_prep(const float* __restrict__ W, int N, int K, int ldw, int Np, int Kp, int transpose,
                           float* __restrict__ W_hi, float* __restrict__ W_lo) {
  // output is [Np][Kp] row-major; source element (n,k) = W[n*ldw + k], or W[k*ldw + n] when transposed
  const int total = Np * Kp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / Kp, k = i % Kp;
    float w = 0.f;
    if (n < N && k < K) w = transpose ? W[(int64_t)k * ldw + n] : W[(int64_t)n * ldw + k];
    const float h = tf32_hi(w);
    W_hi[i] = h;
    W_lo[i] = tf32_hi(w - h);
  }
}

VX_API int vx_mlp_prep(const float* W, int N, int K, int ldw, int Np, int Kp, int transpose, float* W_hi, float* W_lo,
                       cudaStream_t st) {
  VX_REQUIRE(Np % 16 == 0 && Kp % 8 == 0 && Np >= N && Kp >= K, "vx_mlp_prep", "bad padding");
  k_mlp_prep<<<vx_blocks((int64_t)Np * Kp, 256), 256, 0, st>>>(W, N, K, ldw, Np, Kp, transpose, W_hi, W_lo);
  return vx_check_launch("vx_mlp_prep");
}

// ---------------------------------------------------------------------------------------------
// the chain kernel
// ---------------------------------------------------------------------------------------------
struct MlpLayer {
  const float* W_hi;   // [Np][Kp]
  const float* W_lo;
  const float* bias;   // [N] or nullptr
  float* H;            // [cap][ldh] activation output of this layer in HBM (row-major), or nullptr
  const float* mask;   // [cap][ldh]: multiply the output by (mask > 0) (ReLU backward), or nullptr
  float* HT;           // [Np][ldt] the same output TRANSPOSED (feature-major; operand of the split-K weight-gradient GEMM)
  const float* maskT;  // [Np][ldt] transposed mask
  int Kp, Np, N, ldh, relu;
};
struct MlpChain {
  int n_layers;
  MlpLayer L[MLP_MAX_LAYERS];
};

struct __align__(16) MlpSmem {
  float A_lo[MLP_MAXW / 4 * MLP_ROWS * 4];                           // 96 KB  [K/4][128][4]
  float B[MLP_STAGES][2][MLP_SLICE_K / 4 * MLP_MAXW * 4];            // stages x {hi,lo} x [8 chunks][192 rows][4] = 2 x 48 KB
  uint64_t bar_slot[MLP_STAGES];
  uint64_t bar_acc;
  uint32_t tmem_base;
};

// write 8 consecutive features [c0, c0+8) of this thread's row: hi part -> TMEM (A operand), lo part -> smem tile
__device__ __forceinline__ void store_a8(MlpSmem& s, uint32_t tmem_a_lane, int row_in_tile, int c0, const float* v) {
  float hi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) hi[j] = tf32_hi(v[j]);
  tmem_st8(tmem_a_lane + c0, hi);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float4 lo;
    lo.x = tf32_hi(v[4 * q] - hi[4 * q]); lo.y = tf32_hi(v[4 * q + 1] - hi[4 * q + 1]);
    lo.z = tf32_hi(v[4 * q + 2] - hi[4 * q + 2]); lo.w = tf32_hi(v[4 * q + 3] - hi[4 * q + 3]);
    reinterpret_cast<float4*>(s.A_lo)[((c0 >> 2) + q) * MLP_ROWS + row_in_tile] = lo;
  }
}

__global__ void __launch_bounds__(MLP_ROWS, 1)
k_mlp_chain(const float* __restrict__ X, int ldx, int K0, int K0p, const int* __restrict__ n_rows_dev, int capacity, MlpChain ch,
            float* __restrict__ Y, int ldy, int n_out, float* __restrict__ XT, int ldt) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  MlpSmem& s = *reinterpret_cast<MlpSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_rows = min(*n_rows_dev, capacity);
  const int n_tiles = (n_rows + MLP_ROWS - 1) / MLP_ROWS;

  if (tid == 0) {
    for (int i = 0; i < MLP_STAGES; ++i) mbar_init(&s.bar_slot[i], 1);
    mbar_init(&s.bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&s.tmem_base, MLP_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s.tmem_base;
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's 32 TMEM lanes (thread = lane = row)
  uint32_t slot_phase[MLP_STAGES];
#pragma unroll
  for (int i = 0; i < MLP_STAGES; ++i) slot_phase[i] = 0;
  uint32_t acc_phase = 0;
  uint32_t slot_used = 0;   // bit i: slot i has an outstanding commit we must wait for before refilling

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int row = tile * MLP_ROWS + tid;   // this thread's row (global) == TMEM lane
    // ---- stage the input rows: X[row, 0:K0] -> A (hi in TMEM, lo in smem); zeros past n_rows and past K0
    {
      const bool vec = (ldx % 4 == 0);
      const float* src = X + (int64_t)row * ldx;
      for (int c0 = 0; c0 < K0p; c0 += 8) {
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
        if (XT && row < ldt) {   // coalesced: consecutive threads = consecutive rows of one feature line
#pragma unroll
          for (int j = 0; j < 8; ++j) XT[(int64_t)(c0 + j) * ldt + row] = v[j];
        }
        store_a8(s, lane_addr + MLP_TMEM_A, tid, c0, v);
      }
      tmem_st_wait();
    }
    for (int l = 0; l < ch.n_layers; ++l) {
      const MlpLayer& L = ch.L[l];
      const int KS = L.Kp / 8;                                   // K-steps (one tcgen05.mma.kind::tf32 covers K = 8)
      const int NSL = (L.Kp + MLP_SLICE_K - 1) / MLP_SLICE_K;    // weight slices
      const int Np = L.Np;
      const uint32_t idesc = make_idesc_tf32(MLP_ROWS, Np);

      auto load_slice = [&](int sl, int slot) {
        const int kchunks = min(MLP_SLICE_K, L.Kp - sl * MLP_SLICE_K) / 4;   // 16-byte K-chunks in this slice
        const int copies = kchunks * Np;
        for (int c = tid; c < copies; c += MLP_ROWS) {
          const int kc = c / Np, n = c % Np;
          const int64_t g = (int64_t)n * L.Kp + sl * MLP_SLICE_K + kc * 4;
          cp_async16(&s.B[slot][0][(kc * Np + n) * 4], L.W_hi + g);
          cp_async16(&s.B[slot][1][(kc * Np + n) * 4], L.W_lo + g);
        }
        cp_async_commit();
      };

      if (slot_used & 1u) { mbar_wait(&s.bar_slot[0], slot_phase[0]); slot_phase[0] ^= 1; slot_used &= ~1u; }
      load_slice(0, 0);
      for (int sl = 0; sl < NSL; ++sl) {
        const int slot = sl % MLP_STAGES;
        if (sl + 1 < NSL) {
          const int nslot = (sl + 1) % MLP_STAGES;
          if (slot_used & (1u << nslot)) { mbar_wait(&s.bar_slot[nslot], slot_phase[nslot]); slot_phase[nslot] ^= 1; slot_used &= ~(1u << nslot); }
          load_slice(sl + 1, nslot);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        fence_proxy_async();   // cp.async data + this thread's A_lo stores -> visible to the tensor core (async proxy)
        tc_fence_before();     // this thread's tcgen05.st of the A (hi) operand
        __syncthreads();
        if (tid == 0) {
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
          umma_commit(&s.bar_slot[slot]);
          if (sl == NSL - 1) umma_commit(&s.bar_acc);
        }
        slot_used |= (1u << slot);
      }
      // ---- epilogue of layer l: wait for the accumulator, then TMEM -> registers -> (+bias, ReLU / mask) -> next A operand
      mbar_wait(&s.bar_acc, acc_phase);
      acc_phase ^= 1;
      tc_fence_after();
      const bool last = (l == ch.n_layers - 1);
      if (!last) {
        for (int c0 = 0; c0 < Np; c0 += 32) {
          float v[32];
          tmem_ld32(lane_addr + c0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float y = v[j];
            if (L.bias) y += __ldg(L.bias + c0 + j);
            if (L.relu) y = fmaxf(y, 0.f);
            v[j] = y;
          }
          if (L.mask && row < n_rows) {
            const float4* mk = reinterpret_cast<const float4*>(L.mask + (int64_t)row * L.ldh + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 m4 = __ldg(mk + q);
              if (!(m4.x > 0.f)) v[4 * q] = 0.f;
              if (!(m4.y > 0.f)) v[4 * q + 1] = 0.f;
              if (!(m4.z > 0.f)) v[4 * q + 2] = 0.f;
              if (!(m4.w > 0.f)) v[4 * q + 3] = 0.f;
            }
          }
          if (L.maskT && row < n_rows) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (!(__ldg(L.maskT + (int64_t)(c0 + j) * ldt + row) > 0.f)) v[j] = 0.f;
          }
          if (row >= n_rows) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          if (L.HT && row < ldt) {
#pragma unroll
            for (int j = 0; j < 32; ++j) L.HT[(int64_t)(c0 + j) * ldt + row] = v[j];
          }
          if (L.H && row < n_rows) {
            float4* dst = reinterpret_cast<float4*>(L.H + (int64_t)row * L.ldh + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) store_a8(s, lane_addr + MLP_TMEM_A, tid, c0 + 8 * q, v + 8 * q);
        }
        tmem_st_wait();
      } else {
        for (int c0 = 0; c0 < Np; c0 += 16) {
          float v[16];
          tmem_ld16(lane_addr + c0, v);
          if (row < n_rows) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (c0 + j < n_out) {
                float y = v[j];
                if (L.bias) y += __ldg(L.bias + c0 + j);
                Y[(int64_t)row * ldy + c0 + j] = y;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncthreads();   // all TMEM reads / writes + A_lo stores done before the next layer's MMAs
    }
  }
  // drain outstanding slot commits (the barriers must be quiescent before exit)
  for (int i = 0; i < MLP_STAGES; ++i)
    if (slot_used & (1u << i)) mbar_wait(&s.bar_slot[i], slot_phase[i]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, MLP_TMEM_COLS);
}

// Layers are described by two packed HOST arrays so the C ABI stays plain:
//   ptrs_host[l*7 + {0..6}] = device addresses of W_hi, W_lo (prepared, [Np][Kp]), bias, H out, mask, HT out, maskT (0 = none)
//   dims_host[l*5 + {0..4}] = Kp, Np, N, ldh, relu
// X (capacity, ldx) with K0 valid columns; Y (capacity, ldy) receives the first n_out columns of the last layer;
// XT (K0p, ldt) optionally receives the transposed input; all transposed buffers share the row stride ldt >= 128*ceil(capacity/128).
VX_API int vx_mlp_chain(const float* X, int ldx, int K0, const int* n_rows_dev, int capacity, int n_layers,
                        const int64_t* ptrs_host, const int* dims_host, float* Y, int ldy, int n_out, float* XT, int ldt,
                        cudaStream_t st) {
  VX_REQUIRE(n_layers >= 1 && n_layers <= MLP_MAX_LAYERS, "vx_mlp_chain", "1..4 layers");
  VX_REQUIRE(n_rows_dev != nullptr, "vx_mlp_chain", "n_rows_dev required");
  MlpChain ch;
  ch.n_layers = n_layers;
  bool any_t = XT != nullptr;
  for (int l = 0; l < n_layers; ++l) {
    MlpLayer& L = ch.L[l];
    L.W_hi = reinterpret_cast<const float*>(ptrs_host[l * 7 + 0]);
    L.W_lo = reinterpret_cast<const float*>(ptrs_host[l * 7 + 1]);
    L.bias = reinterpret_cast<const float*>(ptrs_host[l * 7 + 2]);
    L.H = reinterpret_cast<float*>(ptrs_host[l * 7 + 3]);
    L.mask = reinterpret_cast<const float*>(ptrs_host[l * 7 + 4]);
    L.HT = reinterpret_cast<float*>(ptrs_host[l * 7 + 5]);
    L.maskT = reinterpret_cast<const float*>(ptrs_host[l * 7 + 6]);
    any_t |= (L.HT != nullptr) || (L.maskT != nullptr);
    L.Kp = dims_host[l * 5 + 0]; L.Np = dims_host[l * 5 + 1]; L.N = dims_host[l * 5 + 2];
    L.ldh = dims_host[l * 5 + 3]; L.relu = dims_host[l * 5 + 4];
    VX_REQUIRE(L.Kp % 8 == 0 && L.Kp >= 8 && L.Kp <= MLP_MAXW && L.Np % 16 == 0 && L.Np >= 16 && L.Np <= MLP_MAXW,
               "vx_mlp_chain", "layer shape");
    if (l + 1 < n_layers)
      VX_REQUIRE(L.Np % 32 == 0 && L.Np == dims_host[(l + 1) * 5 + 0] && (L.ldh % 4 == 0 || (!L.H && !L.mask)),
                 "vx_mlp_chain", "hidden widths must chain, be multiples of 32, and ldh % 4 == 0");
  }
  const int K0p = dims_host[0];
  VX_REQUIRE(K0 <= K0p && K0 <= ldx && n_out <= dims_host[(n_layers - 1) * 5 + 1], "vx_mlp_chain", "K0 / n_out");
  const int tiles_cap = (capacity + MLP_ROWS - 1) / MLP_ROWS;
  VX_REQUIRE(!any_t || ldt >= tiles_cap * MLP_ROWS, "vx_mlp_chain", "ldt must cover whole row tiles");
  static bool attr_set = false;
  const int smem = (int)sizeof(MlpSmem) + 1024;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { vx_set_error("vx_mlp_chain", cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  if (tiles_cap <= 0) return 0;
  const int blocks = min(tiles_cap, vx_num_sms());
  k_mlp_chain<<<blocks, MLP_ROWS, smem, st>>>(X, ldx, K0, K0p, n_rows_dev, capacity, ch, Y, ldy, n_out, XT, ldt);
  return vx_check_launch("vx_mlp_chain");
}

// ---------------------------------------------------------------------------------------------
// Split-K weight-gradient GEMM:  C[m][n] += sum_r At[m][r] * Bt[n][r]   (dW = dY^T H, db = dY^T 1)
// At (M_out, ldt) and Bt (N_in, ldt) are the feature-major ("transposed") activations written by the chain
// kernels; r runs over the MLP rows (K dimension of the MMA, ~43 k), split over the CTAs of blockIdx.x;
// blockIdx.y selects the 128-row M tile.  Both operands are split hi/lo on the fly while being staged (thread-staged
// loads: A hi -> TMEM, A lo / B hi / B lo -> K-major smem tiles, double buffered).  An extra virtual B row of
// ones yields the bias gradient.  Partial results are added to C / c_bias with vector atomics.
// ---------------------------------------------------------------------------------------------
#define DW_KC 32                      // rows (K) per chunk
#define DW_MAXN 208                   // N_in (<= 192) + 1 ones row, padded to 16

struct __align__(16) DwSmem {
  float A_lo[2][DW_KC / 4 * MLP_ROWS * 4];        // 2 x 16 KB
  float B_hi[2][DW_KC / 4 * DW_MAXN * 4];         // 2 x 26 KB
  float B_lo[2][DW_KC / 4 * DW_MAXN * 4];         // 2 x 26 KB
  uint64_t bar_slot[2];
  uint64_t bar_acc;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(MLP_ROWS, 1)
k_mlp_dw(const float* __restrict__ At, int M_out, const float* __restrict__ Bt, int N_in, int ldt,
         const int* __restrict__ n_rows_dev, int capacity, float* __restrict__ C, int ldc, float* __restrict__ c_bias) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  DwSmem& s = *reinterpret_cast<DwSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_rows = min(*n_rows_dev, capacity);
  const int n_chunks = (n_rows + DW_KC - 1) / DW_KC;
  const int m0 = blockIdx.y * MLP_ROWS;
  const int Np = ((N_in + 1) + 15) / 16 * 16;     // + ones row
  if (tid == 0) {
    mbar_init(&s.bar_slot[0], 1); mbar_init(&s.bar_slot[1], 1); mbar_init(&s.bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&s.tmem_base, MLP_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s.tmem_base;
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t idesc = make_idesc_tf32(MLP_ROWS, Np);
  uint32_t phase[2] = {0, 0};
  uint32_t used = 0;
  int it = 0;
  const int m = m0 + tid;                          // this thread's A row (output feature) == TMEM lane
  for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, ++it) {
    const int slot = it & 1;
    const int r0 = chunk * DW_KC;
    // ---- global -> registers (issued before waiting on the slot, to overlap with the in-flight MMAs)
    float a[DW_KC];
    if (m < M_out) {
      const float4* src = reinterpret_cast<const float4*>(At + (int64_t)m * ldt + r0);
#pragma unroll
      for (int q = 0; q < DW_KC / 4; ++q) {
        const float4 x = __ldg(src + q);
        a[4 * q] = x.x; a[4 * q + 1] = x.y; a[4 * q + 2] = x.z; a[4 * q + 3] = x.w;
      }
#pragma unroll
      for (int j = 0; j < DW_KC; ++j)
        if (r0 + j >= n_rows) a[j] = 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < DW_KC; ++j) a[j] = 0.f;
    }
    if (used & (1u << slot)) { mbar_wait(&s.bar_slot[slot], phase[slot]); phase[slot] ^= 1; used &= ~(1u << slot); }
    // ---- A: hi -> TMEM columns [256 + slot*32, +32), lo -> smem
#pragma unroll
    for (int q = 0; q < DW_KC / 8; ++q) {
      float hi[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) hi[j] = tf32_hi(a[8 * q + j]);
      tmem_st8(lane_addr + MLP_TMEM_A + slot * DW_KC + 8 * q, hi);
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        float4 lo;
        lo.x = tf32_hi(a[8 * q + 4 * h2] - hi[4 * h2]); lo.y = tf32_hi(a[8 * q + 4 * h2 + 1] - hi[4 * h2 + 1]);
        lo.z = tf32_hi(a[8 * q + 4 * h2 + 2] - hi[4 * h2 + 2]); lo.w = tf32_hi(a[8 * q + 4 * h2 + 3] - hi[4 * h2 + 3]);
        reinterpret_cast<float4*>(s.A_lo[slot])[(2 * q + h2) * MLP_ROWS + tid] = lo;
      }
    }
    // ---- B: rows n = tid, tid + 128 (< Np); row N_in is the virtual ones row (bias gradient)
    for (int n = tid; n < Np; n += MLP_ROWS) {
#pragma unroll
      for (int q = 0; q < DW_KC / 4; ++q) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < N_in) x = __ldg(reinterpret_cast<const float4*>(Bt + (int64_t)n * ldt + r0) + q);
        else if (n == N_in) x = make_float4(1.f, 1.f, 1.f, 1.f);
        if (r0 + 4 * q + 0 >= n_rows) x.x = 0.f;
        if (r0 + 4 * q + 1 >= n_rows) x.y = 0.f;
        if (r0 + 4 * q + 2 >= n_rows) x.z = 0.f;
        if (r0 + 4 * q + 3 >= n_rows) x.w = 0.f;
        const float4 hi = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
        const float4 lo = make_float4(tf32_hi(x.x - hi.x), tf32_hi(x.y - hi.y), tf32_hi(x.z - hi.z), tf32_hi(x.w - hi.w));
        reinterpret_cast<float4*>(s.B_hi[slot])[q * Np + n] = hi;
        reinterpret_cast<float4*>(s.B_lo[slot])[q * Np + n] = lo;
      }
    }
    tmem_st_wait();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < DW_KC / 8; ++kk) {
        const uint32_t a_tm = tmem + MLP_TMEM_A + slot * DW_KC + kk * 8;
        const uint64_t da_lo = make_desc(smem_u32(s.A_lo[slot]) + (uint32_t)(kk * 2) * MLP_ROWS * 16, MLP_ROWS * 16, 128);
        const uint32_t b_off = (uint32_t)(kk * 2) * Np * 16;
        const uint64_t db_hi = make_desc(smem_u32(s.B_hi[slot]) + b_off, Np * 16, 128);
        const uint64_t db_lo = make_desc(smem_u32(s.B_lo[slot]) + b_off, Np * 16, 128);
        umma_tf32_ts(tmem, a_tm, db_hi, idesc, (it > 0) || (kk > 0));
        umma_tf32_ts(tmem, a_tm, db_lo, idesc, 1);
        umma_tf32_ss(tmem, da_lo, db_hi, idesc, 1);
      }
      umma_commit(&s.bar_slot[slot]);
    }
    used |= (1u << slot);
  }
  if (it > 0) {
    if (tid == 0) umma_commit(&s.bar_acc);
    mbar_wait(&s.bar_acc, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < Np; c0 += 16) {
      float v[16];
      tmem_ld16(lane_addr + c0, v);   // warp-collective: executed by every thread, only the adds are predicated
      if (m < M_out) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int n = c0 + 4 * q;
          if (n + 3 < N_in && (ldc % 4 == 0)) {
            atomicAdd(reinterpret_cast<float4*>(C + (int64_t)m * ldc + n), make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (n + j < N_in) atomicAdd(C + (int64_t)m * ldc + n + j, v[4 * q + j]);
              else if (n + j == N_in && c_bias) atomicAdd(c_bias + m, v[4 * q + j]);
            }
          }
        }
      }
    }
    for (int i = 0; i < 2; ++i)
      if (used & (1u << i)) mbar_wait(&s.bar_slot[i], phase[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, MLP_TMEM_COLS);
}

VX_API int vx_mlp_dw(const float* At, int M_out, const float* Bt, int N_in, int ldt, const int* n_rows_dev, int capacity,
                     float* C, int ldc, float* c_bias, cudaStream_t st) {
  VX_REQUIRE(n_rows_dev != nullptr, "vx_mlp_dw", "n_rows_dev required");
  VX_REQUIRE(M_out >= 1 && M_out <= 256 && N_in >= 1 && N_in + 1 <= DW_MAXN && ldt % 4 == 0, "vx_mlp_dw", "shape");
  static bool attr_set = false;
  const int smem = (int)sizeof(DwSmem) + 1024;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { vx_set_error("vx_mlp_dw", cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  const int m_tiles = (M_out + MLP_ROWS - 1) / MLP_ROWS;
  const int chunks_cap = (capacity + DW_KC - 1) / DW_KC;
  if (chunks_cap <= 0) return 0;
  const int gx = max(1, min(vx_num_sms() / m_tiles, (chunks_cap + 7) / 8));
  k_mlp_dw<<<dim3(gx, m_tiles), MLP_ROWS, smem, st>>>(At, M_out, Bt, N_in, ldt, n_rows_dev, capacity, C, ldc, c_bias);
  return vx_check_launch("vx_mlp_dw");
}
