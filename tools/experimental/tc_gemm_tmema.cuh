// tc_gemm_tmema.cuh -- EXPERIMENT: split-TF32 contraction with the A operand in tensor memory and the B operand as pre-split
// planes by TMA, three-stage ring.
//
//   C[M,N] = A[M,K] B[N,K]^T, dense row-major A, B given as tc_presplit_launch planes; M % 128 == 0, N % 256 == 0, K % 32 == 0.
//
// Why: the library kernel keeps both operands' hi / lo planes in shared memory; per 32-k stage the three MMAs per k-step read
// 144 KB of operands from shared memory and the staging writes 96 KB, 240 KB per stage against 128 B/clk = 1875 cycles, more
// than the 1536 cycles of MMAs: shared-memory bandwidth is what caps it (l1tex is the busiest unit in every ncu capture).
// Here A never touches shared memory: each thread splits its 8 values of one row and writes them to tensor memory
// (tcgen05.st), the MMAs take A from TMEM ([a_tmem] operand form); shared memory only carries B (64 KB TMA write + 96 KB of
// MMA reads per stage = 1250 cycles).  With A out of shared memory a third B stage fits, so a stage's bulk copy is issued
// two stages of MMAs ahead.
#pragma once
#include "../../mvpy_b200/csrc/tc_gemm.cuh"

namespace mvb {
namespace tmema {

using namespace mvb::tc;
constexpr int BNW = 256, NS = 3;
constexpr int B_LBO = BNW * 16 + 16, B_PLANE = KC * B_LBO;
constexpr uint32_t B_BYTES = 2 * B_PLANE;                       // 65 792
constexpr int BAR_OFF = NS * B_BYTES;
constexpr int SMEM_BYTES = BAR_OFF + 256;
constexpr uint32_t COL_A = 256;                                  // accumulator: columns 0..255; A stage s: hi at COL_A + 64 s, lo at + 32
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BNW >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  const uint32_t* u = reinterpret_cast<const uint32_t*>(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]) : "memory");
}

// grid = (N / 256, M / 128)
__global__ void __launch_bounds__(NTHREADS, 1) tmema_kernel(const float* __restrict__ A, const uint8_t* __restrict__ Bplanes,
                                                            float* __restrict__ C, int M, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BNW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);   // [0..2] stage free; [3] chunk done; [4..6] B planes landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 128);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 7; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const int nkb = K / BK;
  const uint8_t* b_tiles = Bplanes + (int64_t)(n0 / BNW) * nkb * (int64_t)B_BYTES;
  const int q = warp & 3, part = warp >> 2;          // TMEM lane quarter; k columns part * 8 .. + 8 of the stage (and accumulator columns part * 64 ..)
  const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
  const float* arow = A + (size_t)(m0 + q * 32 + lane) * K + part * 8;
  float4 a0, a1;
  auto fetch = [&](int kb) {
    a0 = __ldg(reinterpret_cast<const float4*>(arow + kb * BK));
    a1 = __ldg(reinterpret_cast<const float4*>(arow + kb * BK + 4));
  };
  auto load_b = [&](int kb) {     // thread 0 only
    const int sb_ = kb % NS;
    bulk_load(smem_u32(smem + sb_ * B_BYTES), b_tiles + (int64_t)kb * B_BYTES, B_BYTES, smem_u32(&bars[4 + sb_]));
  };
  if (threadIdx.x == 0) { load_b(0); if (nkb > 1) load_b(1); }
  fetch(0);
  float acc[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) acc[j] = 0.f;
  int n_drained = 0;
  for (int it = 0; it < nkb; ++it) {
    const int s = it % NS;
    if (it >= NS) mbar_wait(smem_u32(&bars[s]), (uint32_t)(((it / NS) - 1) & 1));       // stage s of A (TMEM) and B (smem) is free
    {
      float hi[8], lo[8];
      split_tf32(a0.x, hi[0], lo[0]); split_tf32(a0.y, hi[1], lo[1]); split_tf32(a0.z, hi[2], lo[2]); split_tf32(a0.w, hi[3], lo[3]);
      split_tf32(a1.x, hi[4], lo[4]); split_tf32(a1.y, hi[5], lo[5]); split_tf32(a1.z, hi[6], lo[6]); split_tf32(a1.w, hi[7], lo[7]);
      tmem_st8(lane_addr + COL_A + s * 64 + part * 8, hi);
      tmem_st8(lane_addr + COL_A + s * 64 + 32 + part * 8, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    const bool chunk_first = (it % DRAIN) == 0;
    const bool chunk_last = ((it + 1) % DRAIN) == 0 || it == nkb - 1;
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      mbar_wait(smem_u32(&bars[4 + s]), (uint32_t)((it / NS) & 1));
      const uint32_t b_hi = smem_u32(smem + s * B_BYTES), b_lo = b_hi + B_PLANE;
      const uint32_t a_hi = tmem_base + COL_A + s * 64, a_lo = a_hi + 32;
#pragma unroll
      for (int j = 0; j < BK / 8; ++j) {
        const uint64_t db_hi = make_desc(b_hi + 2 * j * B_LBO, B_LBO, 128);
        const uint64_t db_lo = make_desc(b_lo + 2 * j * B_LBO, B_LBO, 128);
        mma_tf32_ts(tmem_base, a_lo + j * 8, db_hi, IDESC, (!chunk_first || j > 0) ? 1u : 0u);
        mma_tf32_ts(tmem_base, a_hi + j * 8, db_lo, IDESC, 1u);
        mma_tf32_ts(tmem_base, a_hi + j * 8, db_hi, IDESC, 1u);
      }
      mma_commit(smem_u32(&bars[s]));
      if (chunk_last) mma_commit(smem_u32(&bars[3]));
      if (it + 2 < nkb) {       // buffer (it + 2) % NS was last read by stage it - 1
        const int s2 = (it + 2) % NS;
        if (it >= 1) mbar_wait(smem_u32(&bars[s2]), (uint32_t)(((it - 1) / NS) & 1));
        load_b(it + 2);
      }
    }
    if (it + 1 < nkb) fetch(it + 1);
    if (chunk_last) {
      mbar_wait(smem_u32(&bars[3]), (uint32_t)(n_drained & 1));
      ++n_drained;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        float v[32];
        tmem_ld32(lane_addr + (uint32_t)(part * 64 + cc * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[cc * 32 + j] += v[j];
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
  }
  {
    const int gm = m0 + q * 32 + lane;
    float* crow = C + (size_t)gm * N + n0 + part * 64;
#pragma unroll
    for (int j = 0; j < 64; j += 4) *reinterpret_cast<float4*>(crow + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

inline cudaError_t tmema_launch(const float* A, const uint8_t* Bplanes, float* C, int M, int N, int K, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tmema_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid(N / 256, M / 128);
  tmema_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(A, Bplanes, C, M, N, K);
  return cudaGetLastError();
}

}  // namespace tmema
}  // namespace mvb
