// sweep_chain.cuh -- the block forward substitution of the LOO sweep as ONE kernel: leverages for all rows x all alphas.
//
//   u_j = [ L_j^-1 | -L_j^-1 W_j ] [ z_j ; u_{j-1} ],   h = sum_j |u_j|^2          (band_loo.cuh, ridgecv.py:230-251)
//
// The recurrence is independent per design row, so a CTA owns one (128-row tile, alpha) pair and walks the whole chain
// j = 0 .. nb-1 by itself: u_{j-1} never leaves the SM.  It lives in shared memory already split into the two TF32
// planes and already in UMMA operand layout, i.e. the epilogue of step j writes the A operand of step j+1.  Per step
// only z_j (128 x 128, shared by the 20 alphas through L2) and the factor block (128 x 256, shared by all row tiles) are
// staged.  Compared with one contraction launch per block this removes the U round trips through HBM (0.8 GB per
// block at BASELINE configs[1]), 100 kernel prologues / epilogues per chain and the per-launch tail.
//
// Same machinery as tc_gemm.cuh: 16 warps stage operands (LDG.128 -> split -> STS.128) into double-buffered K-major
// no-swizzle tiles, one thread issues tcgen05.mma kind::tf32 (3 MMAs per 8-wide k-step), the fp32 accumulator lives in
// TMEM and is folded into registers every 128 k.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tc_gemm.cuh"

namespace mvb {
namespace sc {

constexpr int NB = 128;                 // block size of the tridiagonal factor = tile rows = tile columns
constexpr int SBK = 16;                 // k per stage
constexpr int NTH = 512;
constexpr int LBO = NB * 16 + 16;       // bytes between 16-byte k-chunks of a 128-row operand plane (padded: conflict-free stores)
constexpr int ST_PLANE = (SBK / 4) * LBO;                 // one staged plane: 4 k-chunks
constexpr int STAGE = 4 * ST_PLANE;                        // A hi, A lo, B hi, B lo
constexpr int U_PLANE = (NB / 4) * LBO;                    // resident u_{j-1}: 32 k-chunks
constexpr int OFF_U = 2 * STAGE;
constexpr int OFF_BAR = OFF_U + 2 * U_PLANE;
constexpr int OFF_RED = OFF_BAR + 64;
constexpr int SMEM_BYTES = OFF_RED + 4 * NB * 4;
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// Zblk: [nb][n][128] centred Z = X Q^T in column blocks; MN: [A][nb][128][256] rows [L_j^-1 | -L_j^-1 W_j];
// h: [A][n] = sum_j |u_j|^2.   grid = (A, ceil(n / 128)), 512 threads.
__global__ void __launch_bounds__(NTH, 1) sweep_chain_kernel(const float* __restrict__ Zblk, const float* __restrict__ MN, int64_t n,
                                                             int nb, float* __restrict__ h) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // alpha is the fast grid index: the ~148 CTAs running at a time cover all alphas of a few row tiles, so a z tile is
  // fetched from HBM once and served to the other alphas by L2 (row tiles fastest re-read Z from DRAM once per alpha:
  // ncu measured 40.8 GB of DRAM reads for 2 GB of Z), while the factor blocks of all alphas stay L2 resident
  const int64_t row0 = (int64_t)blockIdx.y * NB;
  const int a = blockIdx.x;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);     // [0],[1]: stage buffer free; [2]: chunk complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 32);
  float* red = reinterpret_cast<float*>(smem + OFF_RED);

  if (threadIdx.x == 0) {
    tc::mbar_init(tc::smem_u32(&bars[0]), 1);
    tc::mbar_init(tc::smem_u32(&bars[1]), 1);
    tc::mbar_init(tc::smem_u32(&bars[2]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "r"((uint32_t)NB)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // u_{-1} = 0
  for (int e = threadIdx.x; e < 2 * U_PLANE / 16; e += NTH) reinterpret_cast<float4*>(smem + OFF_U)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem0 = tc::smem_u32(smem);
  const uint64_t dzero = tc::make_desc(0, LBO, 128);        // descriptor with a zero start address

  // staging map: thread -> (row = tid / 4, 16-byte k-chunk = tid % 4) of the 128 x 16 stage tile
  const int srow = threadIdx.x >> 2, skc = threadIdx.x & 3;
  const bool zrow_ok = row0 + srow < n;
  const float* zsrc = Zblk + (row0 + srow) * NB + skc * 4;                            // + j * n * 128 + k0
  const float* bsrc = MN + ((int64_t)a * nb * NB + srow) * (2 * NB) + skc * 4;        // + j * 128 * 256 + k0
  const uint32_t st_off = (uint32_t)(skc * LBO + srow * 16);

  // accumulator map: thread -> row q*32 + lane, columns part*32 .. +31
  const int q = warp & 3, part = warp >> 2;
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.f;
  float hsum = 0.f;

  constexpr int NSTAGE = 2 * NB / SBK;        // 16 stages per chain step: k = 0..127 from z_j, 128..255 from u_{j-1}
  uint32_t gs = 0;                             // stages issued so far (buffer / barrier parity)
  uint32_t n_folded = 0;
  // register ring: operands of the next PF stages are in flight (a 16-k stage is only ~400 MMA cycles, far less than an
  // L2 round trip, so one stage of lookahead would leave the tensor core waiting on loads)
  constexpr int PF = 4;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 ra[PF], rb[PF];
#pragma unroll
  for (int d = 0; d < PF; ++d) {
    ra[d] = zero4; rb[d] = zero4;
    if (nb > 0) {
      if (zrow_ok) ra[d] = __ldg(reinterpret_cast<const float4*>(zsrc + d * SBK));
      rb[d] = __ldg(reinterpret_cast<const float4*>(bsrc + d * SBK));
    }
  }
  for (int j = 0; j < nb; ++j) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s, ++gs) {
      const float4 va = ra[s % PF], vb = rb[s % PF];
      const int buf = (int)(gs & 1u);
      if (gs >= 2) tc::mbar_wait(tc::smem_u32(&bars[buf]), ((gs >> 1) - 1) & 1u);
      uint8_t* sb = smem + buf * STAGE;
      const bool from_z = s < NSTAGE / 2;
      {
        float4 hi, lo;
        if (from_z) {
          tc::split_tf32(va.x, hi.x, lo.x); tc::split_tf32(va.y, hi.y, lo.y); tc::split_tf32(va.z, hi.z, lo.z); tc::split_tf32(va.w, hi.w, lo.w);
          *reinterpret_cast<float4*>(sb + st_off) = hi;
          *reinterpret_cast<float4*>(sb + ST_PLANE + st_off) = lo;
        }
        tc::split_tf32(vb.x, hi.x, lo.x); tc::split_tf32(vb.y, hi.y, lo.y); tc::split_tf32(vb.z, hi.z, lo.z); tc::split_tf32(vb.w, hi.w, lo.w);
        *reinterpret_cast<float4*>(sb + 2 * ST_PLANE + st_off) = hi;
        *reinterpret_cast<float4*>(sb + 3 * ST_PLANE + st_off) = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      const bool chunk_first = (s % (NSTAGE / 2)) == 0;
      const bool chunk_last = ((s + 1) % (NSTAGE / 2)) == 0;
      if (threadIdx.x == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // descriptors differ only in their 14-bit start-address field (bytes >> 4; all of shared memory fits): one add each
        const uint64_t dstage = dzero + (uint64_t)((smem0 + buf * STAGE) >> 4);
        const uint64_t du = dzero + (uint64_t)((smem0 + OFF_U) >> 4) + (uint64_t)(from_z ? 0 : ((s - NSTAGE / 2) * (SBK / 4) * LBO) >> 4);
#pragma unroll
        for (int ks = 0; ks < SBK / 8; ++ks) {
          constexpr uint64_t KS = (2 * LBO) >> 4;
          const uint64_t da_hi = (from_z ? dstage : du) + ks * KS;
          const uint64_t da_lo = (from_z ? dstage + (ST_PLANE >> 4) : du + (U_PLANE >> 4)) + ks * KS;
          const uint64_t db_hi = dstage + ((2 * ST_PLANE) >> 4) + ks * KS;
          const uint64_t db_lo = dstage + ((3 * ST_PLANE) >> 4) + ks * KS;
          tc::mma_tf32(tmem_base, da_lo, db_hi, IDESC, (!chunk_first || ks > 0) ? 1u : 0u);
          tc::mma_tf32(tmem_base, da_hi, db_lo, IDESC, 1u);
          tc::mma_tf32(tmem_base, da_hi, db_hi, IDESC, 1u);
        }
        tc::mma_commit(tc::smem_u32(&bars[buf]));
        if (chunk_last) tc::mma_commit(tc::smem_u32(&bars[2]));
      }
      // refill the ring slot just consumed with the operands of stage s + PF (possibly of the next chain step)
      {
        const int sn = (s + PF) % NSTAGE;
        const int jn = j + (s + PF) / NSTAGE;
        if (jn < nb) {
          const int k0 = sn * SBK;
          if (sn < NSTAGE / 2) ra[s % PF] = zrow_ok ? __ldg(reinterpret_cast<const float4*>(zsrc + ((int64_t)jn * n) * NB + k0)) : zero4;
          rb[s % PF] = __ldg(reinterpret_cast<const float4*>(bsrc + (int64_t)jn * NB * 2 * NB + k0));
        }
      }
      if (chunk_last) {
        tc::mbar_wait(tc::smem_u32(&bars[2]), n_folded & 1u);
        ++n_folded;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float v[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * 32), v);
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] += v[c];
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
    }
    // u_j complete in registers: leverage contribution, then it becomes the resident A operand of step j + 1
    {
      uint8_t* u_hi = smem + OFF_U;
      uint8_t* u_lo = u_hi + U_PLANE;
      const int r = q * 32 + lane;
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        float4 hi, lo;
        const float x0 = acc[c4 * 4 + 0], x1 = acc[c4 * 4 + 1], x2 = acc[c4 * 4 + 2], x3 = acc[c4 * 4 + 3];
        hsum += x0 * x0 + x1 * x1 + x2 * x2 + x3 * x3;
        tc::split_tf32(x0, hi.x, lo.x); tc::split_tf32(x1, hi.y, lo.y); tc::split_tf32(x2, hi.z, lo.z); tc::split_tf32(x3, hi.w, lo.w);
        const uint32_t off = (uint32_t)((part * 8 + c4) * LBO + r * 16);
        *reinterpret_cast<float4*>(u_hi + off) = hi;
        *reinterpret_cast<float4*>(u_lo + off) = lo;
        acc[c4 * 4 + 0] = 0.f; acc[c4 * 4 + 1] = 0.f; acc[c4 * 4 + 2] = 0.f; acc[c4 * 4 + 3] = 0.f;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // ordered before the next MMAs by the stage barriers
    }
  }
  // h[a][row] = sum over the four column parts
  red[part * NB + q * 32 + lane] = hsum;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < NB && row0 + threadIdx.x < n)
    h[(int64_t)a * n + row0 + threadIdx.x] = red[threadIdx.x] + red[NB + threadIdx.x] + red[2 * NB + threadIdx.x] + red[3 * NB + threadIdx.x];
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)NB) : "memory");
  }
}

inline cudaError_t sweep_chain_launch(const float* Zblk, const float* MN, int64_t n, int nb, int A, float* h, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(sweep_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((unsigned)A, (unsigned)((n + NB - 1) / NB));
  sweep_chain_kernel<<<grid, NTH, SMEM_BYTES, st>>>(Zblk, MN, n, nb, h);
  return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------------------------
// Variant with u_{j-1} resident in TENSOR memory: the accumulator tile (128 columns) and the two TF32 planes of u
// (2 x 128 columns) share one 512-column TMEM allocation; the u-part MMAs take their A operand straight from TMEM
// (tcgen05.mma with a TMEM A operand), so shared memory only holds the streamed operands and can afford 32-k stages
// (half as many block barriers per chain step).
// ---------------------------------------------------------------------------------------------------------------
namespace tm {

constexpr int TBK = 32;
constexpr int T_PLANE = (TBK / 4) * LBO;                    // 8 k-chunks
constexpr int T_STAGE = 4 * T_PLANE;                         // A hi, A lo, B hi, B lo
constexpr int T_OFF_BAR = 2 * T_STAGE;
constexpr int T_OFF_RED = T_OFF_BAR + 64;
constexpr int T_SMEM_BYTES = T_OFF_RED + 4 * NB * 4;
constexpr uint32_t COL_ACC = 0, COL_UHI = 128, COL_ULO = 256, TMEM_COLS = 512;

__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* u = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]), "r"(u[9]),
        "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]), "r"(u[16]), "r"(u[17]), "r"(u[18]), "r"(u[19]),
        "r"(u[20]), "r"(u[21]), "r"(u[22]), "r"(u[23]), "r"(u[24]), "r"(u[25]), "r"(u[26]), "r"(u[27]), "r"(u[28]), "r"(u[29]),
        "r"(u[30]), "r"(u[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(NTH, 1) sweep_chain_tmem_kernel(const float* __restrict__ Zblk, const float* __restrict__ MN, int64_t n,
                                                                  int nb, float* __restrict__ h) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // alpha is the fast grid index: the ~148 CTAs running at a time cover all alphas of a few row tiles, so a z tile is
  // fetched from HBM once and served to the other alphas by L2 (row tiles fastest re-read Z from DRAM once per alpha:
  // ncu measured 40.8 GB of DRAM reads for 2 GB of Z), while the factor blocks of all alphas stay L2 resident
  const int64_t row0 = (int64_t)blockIdx.y * NB;
  const int a = blockIdx.x;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T_OFF_BAR);   // [0],[1]: stage buffer free; [2]: chunk complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + T_OFF_BAR + 32);
  float* red = reinterpret_cast<float*>(smem + T_OFF_RED);

  if (threadIdx.x == 0) {
    tc::mbar_init(tc::smem_u32(&bars[0]), 1);
    tc::mbar_init(tc::smem_u32(&bars[1]), 1);
    tc::mbar_init(tc::smem_u32(&bars[2]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem0 = tc::smem_u32(smem);
  const uint64_t dzero = tc::make_desc(0, LBO, 128);

  // accumulator / u map: thread -> TMEM lane (row) q*32 + lane, columns part*32 .. +31
  const int q = warp & 3, part = warp >> 2;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.f;
  tmem_st32(lane_addr + COL_UHI + part * 32, acc);          // u_{-1} = 0
  tmem_st32(lane_addr + COL_ULO + part * 32, acc);
  tmem_st_wait();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float hsum = 0.f;

  // staging map: thread -> row tid / 4, 16-byte k-chunks (tid % 4) and (tid % 4) + 4 of the 128 x 32 stage tile
  const int srow = threadIdx.x >> 2, skc = threadIdx.x & 3;
  const bool zrow_ok = row0 + srow < n;
  const float* zsrc = Zblk + (row0 + srow) * NB + skc * 4;
  const float* bsrc = MN + ((int64_t)a * nb * NB + srow) * (2 * NB) + skc * 4;
  const uint32_t st_off0 = (uint32_t)(skc * LBO + srow * 16), st_off1 = (uint32_t)((skc + 4) * LBO + srow * 16);

  constexpr int NSTAGE = 2 * NB / TBK;        // 8 stages per chain step
  constexpr int PF = 3;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 ra0[PF], ra1[PF], rb0[PF], rb1[PF];
#pragma unroll
  for (int d = 0; d < PF; ++d) {
    ra0[d] = zero4; ra1[d] = zero4; rb0[d] = zero4; rb1[d] = zero4;
    if (nb > 0) {
      if (zrow_ok) { ra0[d] = __ldg(reinterpret_cast<const float4*>(zsrc + d * TBK)); ra1[d] = __ldg(reinterpret_cast<const float4*>(zsrc + d * TBK + 16)); }
      rb0[d] = __ldg(reinterpret_cast<const float4*>(bsrc + d * TBK)); rb1[d] = __ldg(reinterpret_cast<const float4*>(bsrc + d * TBK + 16));
    }
  }
  uint32_t gs = 0, n_folded = 0;
  auto put = [](uint8_t* hi_p, uint8_t* lo_p, uint32_t off, const float4& v) {
    float4 hi, lo;
    tc::split_tf32(v.x, hi.x, lo.x); tc::split_tf32(v.y, hi.y, lo.y); tc::split_tf32(v.z, hi.z, lo.z); tc::split_tf32(v.w, hi.w, lo.w);
    *reinterpret_cast<float4*>(hi_p + off) = hi;
    *reinterpret_cast<float4*>(lo_p + off) = lo;
  };
  for (int j = 0; j < nb; ++j) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s, ++gs) {
      const int buf = (int)(gs & 1u);
      if (gs >= 2) tc::mbar_wait(tc::smem_u32(&bars[buf]), ((gs >> 1) - 1) & 1u);
      uint8_t* sb = smem + buf * T_STAGE;
      const bool from_z = s < NSTAGE / 2;
      if (from_z) { put(sb, sb + T_PLANE, st_off0, ra0[s % PF]); put(sb, sb + T_PLANE, st_off1, ra1[s % PF]); }
      put(sb + 2 * T_PLANE, sb + 3 * T_PLANE, st_off0, rb0[s % PF]);
      put(sb + 2 * T_PLANE, sb + 3 * T_PLANE, st_off1, rb1[s % PF]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      const bool chunk_first = (s % (NSTAGE / 2)) == 0;
      const bool chunk_last = ((s + 1) % (NSTAGE / 2)) == 0;
      if (threadIdx.x == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t dstage = dzero + (uint64_t)((smem0 + buf * T_STAGE) >> 4);
#pragma unroll
        for (int ks = 0; ks < TBK / 8; ++ks) {
          constexpr uint64_t KS = (2 * LBO) >> 4;
          const uint64_t db_hi = dstage + ((2 * T_PLANE) >> 4) + ks * KS;
          const uint64_t db_lo = dstage + ((3 * T_PLANE) >> 4) + ks * KS;
          const uint32_t accum = (!chunk_first || ks > 0) ? 1u : 0u;
          if (from_z) {
            const uint64_t da_hi = dstage + ks * KS, da_lo = dstage + (T_PLANE >> 4) + ks * KS;
            tc::mma_tf32(tmem_base + COL_ACC, da_lo, db_hi, IDESC, accum);
            tc::mma_tf32(tmem_base + COL_ACC, da_hi, db_lo, IDESC, 1u);
            tc::mma_tf32(tmem_base + COL_ACC, da_hi, db_hi, IDESC, 1u);
          } else {
            const uint32_t kcol = (uint32_t)((s - NSTAGE / 2) * TBK + ks * 8);     // k index inside u_{j-1}
            mma_tf32_ts(tmem_base + COL_ACC, tmem_base + COL_ULO + kcol, db_hi, IDESC, accum);
            mma_tf32_ts(tmem_base + COL_ACC, tmem_base + COL_UHI + kcol, db_lo, IDESC, 1u);
            mma_tf32_ts(tmem_base + COL_ACC, tmem_base + COL_UHI + kcol, db_hi, IDESC, 1u);
          }
        }
        tc::mma_commit(tc::smem_u32(&bars[buf]));
        if (chunk_last) tc::mma_commit(tc::smem_u32(&bars[2]));
      }
      {
        const int sn = (s + PF) % NSTAGE;
        const int jn = j + (s + PF) / NSTAGE;
        if (jn < nb) {
          const int k0 = sn * TBK;
          if (sn < NSTAGE / 2) {
            const float* zp = zsrc + ((int64_t)jn * n) * NB + k0;
            ra0[s % PF] = zrow_ok ? __ldg(reinterpret_cast<const float4*>(zp)) : zero4;
            ra1[s % PF] = zrow_ok ? __ldg(reinterpret_cast<const float4*>(zp + 16)) : zero4;
          }
          const float* bp = bsrc + (int64_t)jn * NB * 2 * NB + k0;
          rb0[s % PF] = __ldg(reinterpret_cast<const float4*>(bp));
          rb1[s % PF] = __ldg(reinterpret_cast<const float4*>(bp + 16));
        }
      }
      if (chunk_last) {
        tc::mbar_wait(tc::smem_u32(&bars[2]), n_folded & 1u);
        ++n_folded;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float v[32];
        tc::tmem_ld32(lane_addr + COL_ACC + (uint32_t)(part * 32), v);
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] += v[c];
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
    }
    // u_j complete: leverage contribution, then both TF32 planes go back to TMEM as the A operand of step j + 1
    {
      float hi[32], lo[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        hsum += acc[c] * acc[c];
        tc::split_tf32(acc[c], hi[c], lo[c]);
        acc[c] = 0.f;
      }
      tmem_st32(lane_addr + COL_UHI + part * 32, hi);
      tmem_st32(lane_addr + COL_ULO + part * 32, lo);
      tmem_st_wait();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // ordered before the next MMAs by the stage barriers
    }
  }
  red[part * NB + q * 32 + lane] = hsum;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < NB && row0 + threadIdx.x < n)
    h[(int64_t)a * n + row0 + threadIdx.x] = red[threadIdx.x] + red[NB + threadIdx.x] + red[2 * NB + threadIdx.x] + red[3 * NB + threadIdx.x];
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace tm

inline cudaError_t sweep_chain_tmem_launch(const float* Zblk, const float* MN, int64_t n, int nb, int A, float* h, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tm::sweep_chain_tmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tm::T_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((unsigned)A, (unsigned)((n + NB - 1) / NB));
  tm::sweep_chain_tmem_kernel<<<grid, NTH, tm::T_SMEM_BYTES, st>>>(Zblk, MN, n, nb, h);
  return cudaGetLastError();
}

}  // namespace sc
}  // namespace mvb
