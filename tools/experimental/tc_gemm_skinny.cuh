// tc_gemm_skinny.cuh -- EXPERIMENT, OFF BY DEFAULT (not measured yet: written at the end of r01 after the GPU budget was spent).
// Standalone check: `make -C tools tc_gemm_skinny_test`, run on a B200 first thing in r02; in the library it is used by the
// block-Lanczos reduction only when MVB_LB_SKINNY=1 is set (lanczos_band.cuh, gemm()).
//
//   W[128][N] = Q[128][K] * G[N][K]^T        (the block-Lanczos products: one 128-row block against a large dense matrix)
//
// In the library these run as M = 128 products with G pre-split into planes as the B operand: one M tile, so every CTA streams its
// own share of the 8 N K bytes of planes (1.34 GB at cfg2) from HBM through a single cp.async.bulk stream (~16 B/clk per SM
// measured: 0.41 ms per product, 104 TFLOP/s, the tensor pipe waits).  Here the product is turned around: the rows of G are the
// MMA's A operand (128 rows per CTA, N / 128 CTAs x split-K), read as raw fp32 (4 N K bytes, 0.67 GB) by the 512 threads with
// PF stages of loads in flight per thread, split in registers and written to tensor memory; the 128-row block Q is the B operand,
// pre-split into planes (8 * 128 * K bytes = 13 MB, L2-resident) and delivered by TMA into a four-stage ring; the accumulator
// tile (rows = rows of G) is stored transposed so that the output keeps the [128][N] layout the reduction expects.
// Expected: HBM needs 0.67 GB / ~6.5 TB/s = 0.10 ms per product; the tensor pipe at the TF32 rate cuBLAS reaches here (738 TFLOP/s,
// i.e. ~103 cycles per 128 x 128 x 8 MMA, 12 per stage) needs ~0.18 ms with all SMs busy -- against 0.41 ms measured now.
// Open question for the measurement: each lane reads its own row (32 B per row per warp load, the four column parts of a stage
// make up one 128-byte line from four warps); if HBM efficiency suffers, let warp group g own whole stages it = g (mod 4)
// (one 128-byte line per lane, tcgen05.st x32) with a per-group mbarrier instead of the block barrier.
//
// Integration (lanczos_band.cuh, `gemm()`: every call there has M = NB = 128): W = Q_j Gp (N = K = P), H = W Qt^T (N = nq,
// K = P) and W -= H Q (N = P, K = nq; via the partial-sum buffer with subtract) all have a dense row-major B with ld = P.
// With MVB_LB_SKINNY=1 the large ones pre-split their 128-row A block per call (13 MB) into the scratch that otherwise holds
// the Gp planes and launch skinny_kernel with the split count from tc_choose_splits(tiles = N / 128, ...); reduce_parts_kernel
// is unchanged (the transposed store writes the same [split][128][N] layout).  Once measured faster, the Gp / Qt / Q^T plane
// copies (3 x 1.34 GB and their incremental presplit launches) can go.
#pragma once
#include "../../mvpy_b200/csrc/tc_gemm.cuh"

namespace mvb {
namespace skinny {

using namespace mvb::tc;
constexpr int BNQ = 128;                                     // rows of the block (MMA N)
constexpr int NS = 4;                                        // ring depth: B stages in shared memory, A stages in TMEM, loads in flight
static_assert(NS == DRAIN, "the unrolled stage loop assumes ring depth == drain period");
constexpr int Q_LBO = BNQ * 16 + 16, Q_PLANE = KC * Q_LBO;   // 2064, 16 512
constexpr uint32_t Q_BYTES = 2 * Q_PLANE;                    // 33 024 per (128-row tile, 32-k stage)
constexpr int BAR_OFF = NS * (int)Q_BYTES;                   // [0..3] stage free; [4] chunk done; [5..8] B planes landed
constexpr int SMEM_BYTES = BAR_OFF + 128;
constexpr uint32_t COL_A = 128, TMEM_COLS = 512;             // accumulator: columns 0..127; A stage s: hi at COL_A + 64 s, lo at + 32
constexpr int NACC = BNQ / NPART;                            // 32 accumulator columns per thread

// Q (128 x K, row-major, ld) -> planes[kb][plane][kc * Q_LBO + r * 16 + (k % 4) * 4]; K % 32 == 0
__global__ void presplit128_kernel(const float* __restrict__ Q, int64_t ld, int K, uint8_t* __restrict__ planes,
                                   const unsigned int* __restrict__ skip_flag) {
  if (skip_flag && *skip_flag) return;
  const int64_t total = (int64_t)BNQ * (K / 4);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(e % (K / 4)), r = (int)(e / (K / 4));      // consecutive threads: consecutive 16-byte chunks of one row
    const int kb = c4 / KC, kc = c4 % KC;
    const float4 v = __ldg(reinterpret_cast<const float4*>(Q + (int64_t)r * ld + c4 * 4));
    float4 h, l;
    split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
    uint8_t* dst = planes + (int64_t)kb * Q_BYTES + kc * Q_LBO + r * 16;
    *reinterpret_cast<float4*>(dst) = h;
    *reinterpret_cast<float4*>(dst + Q_PLANE) = l;
    if (r == 0) {                                                   // the 16 bytes of padding that end every chunk
      *reinterpret_cast<float4*>(dst + BNQ * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(dst + Q_PLANE + BNQ * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// grid = (N / 128, splits); W + split * w_split_stride receives the partial sum of the split's K range
__global__ void __launch_bounds__(NTHREADS, 1) skinny_kernel(const float* __restrict__ G, int64_t ldg, const uint8_t* __restrict__ Qplanes,
                                                             float* __restrict__ W, int64_t ldw, int64_t w_split_stride, int N, int K,
                                                             const unsigned int* __restrict__ skip_flag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (skip_flag && *skip_flag) return;          // the whole launch is a no-op (conditional clean-up rounds)
  const int n0 = blockIdx.x * BM;
  const int splits = gridDim.y, split_id = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 96);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * NS + 1; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const int nkb_total = K / BK;
  const int kb_per = (nkb_total + splits - 1) / splits;
  const int kb_begin = split_id * kb_per;
  const int nkb = max(0, min(nkb_total, kb_begin + kb_per) - kb_begin);

  const uint8_t* q_tiles = Qplanes + (int64_t)kb_begin * Q_BYTES;
  const int q = warp & 3, part = warp >> 2;     // TMEM lane quarter; k columns part * 8 .. + 8 of a stage, accumulator columns part * 32 ..
  const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
  const float* grow = G + (int64_t)(n0 + q * 32 + lane) * ldg + (int64_t)kb_begin * BK + part * 8;
  float4 a0[NS], a1[NS];                        // NS stages of this thread's 8 values in flight (statically indexed: loop unrolled by NS)
  auto load_b = [&](int it) {                   // thread 0 only
    const int sb = it % NS;
    bulk_load(smem_u32(smem + sb * Q_BYTES), q_tiles + (int64_t)it * Q_BYTES, Q_BYTES, smem_u32(&bars[NS + 1 + sb]));
  };
  if (threadIdx.x == 0)
    for (int i = 0; i < NS - 1 && i < nkb; ++i) load_b(i);
#pragma unroll
  for (int u = 0; u < NS; ++u)
    if (u < nkb) {
      a0[u] = __ldg(reinterpret_cast<const float4*>(grow + u * BK));
      a1[u] = __ldg(reinterpret_cast<const float4*>(grow + u * BK + 4));
    }
  float acc[NACC];
#pragma unroll
  for (int j = 0; j < NACC; ++j) acc[j] = 0.f;
  int n_drained = 0;

  for (int it0 = 0; it0 < nkb; it0 += NS) {
#pragma unroll
    for (int u = 0; u < NS; ++u) {
      const int it = it0 + u;
      if (it < nkb) {                                                                    // block-uniform
        if (it >= NS) mbar_wait(smem_u32(&bars[u]), (uint32_t)(((it / NS) - 1) & 1));    // stage u (A in TMEM, B in smem) is free
        {
          float hi[8], lo[8];
          split_tf32(a0[u].x, hi[0], lo[0]); split_tf32(a0[u].y, hi[1], lo[1]); split_tf32(a0[u].z, hi[2], lo[2]); split_tf32(a0[u].w, hi[3], lo[3]);
          split_tf32(a1[u].x, hi[4], lo[4]); split_tf32(a1[u].y, hi[5], lo[5]); split_tf32(a1[u].z, hi[6], lo[6]); split_tf32(a1[u].w, hi[7], lo[7]);
          tmem_st8(lane_addr + COL_A + u * 64 + part * 8, hi);
          tmem_st8(lane_addr + COL_A + u * 64 + 32 + part * 8, lo);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        if (it + NS < nkb) {                     // the slot is consumed: its next load (NS stages ahead) leaves at once
          a0[u] = __ldg(reinterpret_cast<const float4*>(grow + (int64_t)(it + NS) * BK));
          a1[u] = __ldg(reinterpret_cast<const float4*>(grow + (int64_t)(it + NS) * BK + 4));
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        const bool chunk_last = (u == NS - 1) || it == nkb - 1;
        if (threadIdx.x == 0) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          mbar_wait(smem_u32(&bars[NS + 1 + u]), (uint32_t)((it / NS) & 1));             // this stage's Q planes have landed
          const uint32_t b_hi = smem_u32(smem + u * Q_BYTES), b_lo = b_hi + Q_PLANE;
          const uint32_t a_hi = tmem_base + COL_A + u * 64, a_lo = a_hi + 32;
#pragma unroll
          for (int j = 0; j < BK / 8; ++j) {
            const uint64_t db_hi = make_desc(b_hi + 2 * j * Q_LBO, Q_LBO, 128);
            const uint64_t db_lo = make_desc(b_lo + 2 * j * Q_LBO, Q_LBO, 128);
            mma_tf32_ta(tmem_base, a_lo + j * 8, db_hi, Cfg<BNQ>::IDESC, (u > 0 || j > 0) ? 1u : 0u);
            mma_tf32_ta(tmem_base, a_hi + j * 8, db_lo, Cfg<BNQ>::IDESC, 1u);
            mma_tf32_ta(tmem_base, a_hi + j * 8, db_hi, Cfg<BNQ>::IDESC, 1u);
          }
          mma_commit(smem_u32(&bars[u]));
          if (chunk_last) mma_commit(smem_u32(&bars[NS]));
          if (it + NS - 1 < nkb) {               // buffer (it + NS - 1) % NS was last read by stage it - 1
            const int s2 = (u + NS - 1) % NS;
            if (it >= 1) mbar_wait(smem_u32(&bars[s2]), (uint32_t)(((it - 1) / NS) & 1));
            load_b(it + NS - 1);
          }
        }
        if (chunk_last) {
          mbar_wait(smem_u32(&bars[NS]), (uint32_t)(n_drained & 1));
          ++n_drained;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          float v[32];
          tmem_ld32(lane_addr + (uint32_t)(part * NACC), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] += v[j];
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
      }
    }
  }
  {   // transposed store: this thread holds row gn of G x block rows part * 32 ..; lanes sit on consecutive gn: 128-byte rows of W
    const int gn = n0 + q * 32 + lane;
    float* w = W + (int64_t)split_id * w_split_stride + gn;
    if (gn < N) {
#pragma unroll
      for (int j = 0; j < NACC; ++j) w[(int64_t)(part * NACC + j) * ldw] = acc[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

inline size_t planes128_bytes(int K) { return (size_t)(K / BK) * Q_BYTES; }

inline cudaError_t presplit128_launch(const float* Q, int64_t ld, int K, uint8_t* planes, cudaStream_t st,
                                      const unsigned int* skip_flag = nullptr) {
  if (K % BK) return cudaErrorInvalidValue;
  presplit128_kernel<<<148, 256, 0, st>>>(Q, ld, K, planes, skip_flag);
  return cudaGetLastError();
}

// W[split][128][ldw] partial sums; N % 128 == 0, K % 32 == 0
inline cudaError_t skinny_launch(const float* G, int64_t ldg, const uint8_t* Qplanes, float* W, int64_t ldw, int64_t w_split_stride,
                                 int N, int K, int splits, cudaStream_t st, const unsigned int* skip_flag = nullptr) {
  if (N % BM || K % BK || splits < 1) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return e;
  skinny_kernel<<<dim3(N / BM, splits), NTHREADS, SMEM_BYTES, st>>>(G, ldg, Qplanes, W, ldw, w_split_stride, N, K, skip_flag);
  return cudaGetLastError();
}

}  // namespace skinny
}  // namespace mvb
