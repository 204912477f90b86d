// tc_gemm_tmema2.cuh -- EXPERIMENT (not measured yet: written at the end of r01 after the GPU budget was spent; build with
// `make -C tools tc_gemm_tmema2_test`, run on a B200 in r02).  tc_gemm_tmema.cuh with the accumulator fold taken off the MMA
// critical path.
//
//   C[M,N] = A[M,K] B[N,K]^T, dense row-major A, B given as tc_presplit_launch planes; M % 128 == 0, N % 256 == 0, K % 32 == 0.
//
// In tc_gemm_ta_kernel the tensor pipe stops every 128 k while all threads read the accumulator tile out of TMEM (the fold that
// keeps the truncating tensor-core accumulation short), ~13 % of the time by the ncu tensor-active figure against cuBLAS TF32.
// Here the 256-wide tile is accumulated as two 128-wide halves whose 128-k chunks are staggered by two stages, over three
// 128-column TMEM buffers: when a half finishes a chunk it moves on to the buffer that was drained two stages earlier, and the
// finished buffer is read by the threads while the MMAs of the next stages run.  TMEM: 3 x 128 accumulator columns + 2 A stages
// x 64 = 512.  Left half chunk c = stages 4c .. 4c+3 in buffer c % 3; right half chunk d = stages 4d-2 .. 4d+1 in buffer
// (d + 1) % 3.  One commit barrier per stage (ring of four) serves the A-stage reuse, the B-stage reuse and the drains.
#pragma once
#include "../../mvpy_b200/csrc/tc_gemm.cuh"

namespace mvb {
namespace tmema2 {

using namespace mvb::tc;
constexpr int BNW = 256, NSB = 3, NSA = 2;
constexpr int B_LBO = BNW * 16 + 16, B_PLANE = KC * B_LBO;
constexpr uint32_t B_BYTES = 2 * B_PLANE;                       // 65 792
constexpr int BAR_OFF = NSB * B_BYTES;                           // [0..3] MMAs of stage it % 4 done; [4..6] B planes landed
constexpr int SMEM_BYTES = BAR_OFF + 128;
constexpr uint32_t COL_A = 384;                                  // accumulator buffers at columns 0, 128, 256; A stage s: hi at COL_A + 64 s, lo at + 32
constexpr uint32_t IDESC128 = Cfg<128>::IDESC;

__device__ __forceinline__ void wait_stage_done(uint64_t* bars, int stage) {      // MMAs of stage `stage` (and all before) are complete
  mbar_wait(smem_u32(&bars[stage & 3]), (uint32_t)((stage >> 2) & 1));
}

// grid = (N / 256, M / 128)
__global__ void __launch_bounds__(NTHREADS, 1) tmema2_kernel(const float* __restrict__ A, const uint8_t* __restrict__ Bplanes,
                                                             float* __restrict__ C, int M, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BNW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 96);
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
  const int q = warp & 3, part = warp >> 2;          // TMEM lane quarter; k columns part * 8 .. + 8 of a stage; 32 accumulator columns of each half
  const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
  const float* arow = A + (size_t)(m0 + q * 32 + lane) * K + part * 8;
  float4 a0, a1;
  auto fetch = [&](int kb) {
    a0 = __ldg(reinterpret_cast<const float4*>(arow + kb * BK));
    a1 = __ldg(reinterpret_cast<const float4*>(arow + kb * BK + 4));
  };
  auto load_b = [&](int kb) {     // thread 0 only
    const int sb = kb % NSB;
    bulk_load(smem_u32(smem + sb * B_BYTES), b_tiles + (int64_t)kb * B_BYTES, B_BYTES, smem_u32(&bars[4 + sb]));
  };
  float accL[32], accR[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) { accL[j] = 0.f; accR[j] = 0.f; }
  auto drain = [&](float* acc, int buf) {            // buf's MMAs are complete (caller waited): fold the half tile into registers
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float v[32];
    tmem_ld32(lane_addr + (uint32_t)(buf * 128 + part * 32), v);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] += v[j];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  };
  if (nkb > 0) {
    if (threadIdx.x == 0) { load_b(0); if (nkb > 1) load_b(1); }
    fetch(0);
  }
  for (int it = 0; it < nkb; ++it) {
    const int sa = it % NSA, sb = it % NSB;
    if (it >= NSA) wait_stage_done(bars, it - NSA);                       // A stage sa in TMEM is free
    {
      float hi[8], lo[8];
      split_tf32(a0.x, hi[0], lo[0]); split_tf32(a0.y, hi[1], lo[1]); split_tf32(a0.z, hi[2], lo[2]); split_tf32(a0.w, hi[3], lo[3]);
      split_tf32(a1.x, hi[4], lo[4]); split_tf32(a1.y, hi[5], lo[5]); split_tf32(a1.z, hi[6], lo[6]); split_tf32(a1.w, hi[7], lo[7]);
      tmem_st8(lane_addr + COL_A + sa * 64 + part * 8, hi);
      tmem_st8(lane_addr + COL_A + sa * 64 + 32 + part * 8, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      mbar_wait(smem_u32(&bars[4 + sb]), (uint32_t)((it / NSB) & 1));     // this stage's B planes have landed
      const uint32_t b_hi = smem_u32(smem + sb * B_BYTES), b_lo = b_hi + B_PLANE;
      const uint32_t a_hi = tmem_base + COL_A + sa * 64, a_lo = a_hi + 32;
      const int bufL = (it >> 2) % 3, bufR = (((it + 2) >> 2) + 1) % 3;
      const bool firstL = (it & 3) == 0, firstR = ((it + 2) & 3) == 0 || it == 0;
#pragma unroll
      for (int j = 0; j < BK / 8; ++j) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t d = tmem_base + (uint32_t)((half ? bufR : bufL) * 128);
          const bool first = half ? firstR : firstL;
          const uint64_t db_hi = make_desc(b_hi + 2 * j * B_LBO + half * 128 * 16, B_LBO, 128);
          const uint64_t db_lo = make_desc(b_lo + 2 * j * B_LBO + half * 128 * 16, B_LBO, 128);
          mma_tf32_ta(d, a_lo + j * 8, db_hi, IDESC128, (!first || j > 0) ? 1u : 0u);
          mma_tf32_ta(d, a_hi + j * 8, db_lo, IDESC128, 1u);
          mma_tf32_ta(d, a_hi + j * 8, db_hi, IDESC128, 1u);
        }
      }
      mma_commit(smem_u32(&bars[it & 3]));
      if (it + 2 < nkb) {       // B buffer (it + 2) % 3 was last read by stage it - 1
        if (it >= 1) wait_stage_done(bars, it - 1);
        load_b(it + 2);
      }
    }
    if (it + 1 < nkb) fetch(it + 1);
    // a half whose chunk ended with stage it - 1 is folded now, while the MMAs of stage it run on the other buffers
    if (it >= 4 && (it & 3) == 0) { wait_stage_done(bars, it - 1); drain(accL, ((it - 1) >> 2) % 3); }
    if (it >= 2 && (it & 3) == 2) { wait_stage_done(bars, it - 1); drain(accR, (((it + 1) >> 2) + 1) % 3); }
  }
  if (nkb > 0) {                // the chunks that contain the last stage
    wait_stage_done(bars, nkb - 1);
    drain(accL, ((nkb - 1) >> 2) % 3);
    drain(accR, (((nkb + 1) >> 2) + 1) % 3);
  }
  {
    const int gm = m0 + q * 32 + lane;
    float* crow = C + (size_t)gm * N + n0 + part * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      *reinterpret_cast<float4*>(crow + j) = make_float4(accL[j], accL[j + 1], accL[j + 2], accL[j + 3]);
      *reinterpret_cast<float4*>(crow + 128 + j) = make_float4(accR[j], accR[j + 1], accR[j + 2], accR[j + 3]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

inline cudaError_t tmema2_launch(const float* A, const uint8_t* Bplanes, float* C, int M, int N, int K, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(tmema2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return e;
  dim3 grid(N / 256, M / 128);
  tmema2_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(A, Bplanes, C, M, N, K);
  return cudaGetLastError();
}

}  // namespace tmema2
}  // namespace mvb
