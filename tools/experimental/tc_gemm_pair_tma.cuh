// tc_gemm_pair_tma.cuh -- EXPERIMENT: CTA pairs (tcgen05 cta_group::2) + the dense B operand as pre-split TF32 planes by TMA.
//
// As tc_gemm_pair.cuh, but CTA r's half of the B tile (128 rows x 32 k, hi + lo planes = 33 024 B, the shared-memory image
// written by presplit128_kernel) arrives by one cp.async.bulk per stage, and the ring has three stages: the copy of stage
// it + 2 is issued when stage it - 1 has been consumed and has two stages of MMAs to land.  A is staged by the threads.
//
//   C[M,N] = A[M,K] B[N,K]^T, dense row-major operands, M % 256 == 0, N % 256 == 0, K % 32 == 0.
//
// A cluster of two CTAs owns a 256 x 256 output tile.  CTA r stages rows [128 r, 128 r + 128) of the A tile and rows
// [128 r, 128 r + 128) of the B tile (half of what a lone CTA stages for a 128 x 256 tile: the staging loop is the
// bottleneck of tc_gemm.cuh); the leader issues tcgen05.mma.cta_group::2 (M = 256, N = 256), each SM's tensor core
// produces its 128 accumulator rows in its own TMEM, both CTAs fold and store their halves.
// Synchronisation: per stage a "full" mbarrier in the leader (one arrival per CTA: local, and remote through
// mapa + mbarrier.arrive.shared::cluster), "free" / "chunk done" barriers in both CTAs through multicast tcgen05.commit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace mvb {
namespace pairtma {

constexpr int BM = 128, BNH = 128, BN = 256, BK = 32, KC = BK / 4, NTHREADS = 512, DRAIN = 4;
constexpr int LBO = 128 * 16 + 16;
constexpr int PLANE = KC * LBO;
constexpr int STAGE = 4 * PLANE;                 // A hi, A lo, B-half hi, B-half lo
constexpr int NS = 3;
constexpr int BAR_OFF = NS * STAGE;
constexpr int SMEM_BYTES = BAR_OFF + 256;
constexpr uint32_t BHALF_BYTES = 2 * PLANE;        // this CTA's B half of one stage: hi plane + lo plane
// D = F32, A = B = TF32, K-major, N = 256, M = 256
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t it = 0; it < (1u << 24); ++it) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
  }
  printf("pair gemm: mbarrier timeout block %d thread %d bar %u\n", blockIdx.x, threadIdx.x, bar);
  __trap();
}
// arrive on the barrier at the same shared-memory offset in CTA `target` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t local_bar, uint32_t target) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(target));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void mma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void commit2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
                 "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]),
                 "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]),
                 "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void split(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}

__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// B [N][K] row-major -> planes[(tile128 * nkb + kb) * BHALF_BYTES + plane * PLANE + kc * LBO + r * 16]
__global__ void presplit128_kernel(const float* __restrict__ B, int N, int K, uint8_t* __restrict__ planes) {
  const int nkb = K / BK;
  const int64_t total = (int64_t)N * nkb * KC;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int kc = (int)(e % KC);
    const int64_t t = e / KC;
    const int row = (int)(t % N), kb = (int)(t / N);
    const float4 v = __ldg(reinterpret_cast<const float4*>(B + (size_t)row * K + kb * BK + kc * 4));
    float4 h, l;
    split(v.x, h.x, l.x); split(v.y, h.y, l.y); split(v.z, h.z, l.z); split(v.w, h.w, l.w);
    uint8_t* dst = planes + ((int64_t)(row / 128) * nkb + kb) * BHALF_BYTES + kc * LBO + (row % 128) * 16;
    *reinterpret_cast<float4*>(dst) = h;
    *reinterpret_cast<float4*>(dst + PLANE) = l;
    if (row % 128 == 0) {
      *reinterpret_cast<float4*>(dst + 128 * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(dst + PLANE + 128 * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// grid = (2 * M / 256, N / 256), cluster (2, 1, 1)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
pair_tma_kernel(const float* __restrict__ A, const uint8_t* __restrict__ Bplanes, float* __restrict__ C, int M, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t rank = cluster_rank();
  const int m0 = (blockIdx.x >> 1) * 256 + (int)rank * BM;          // this CTA's A rows = its accumulator rows
  const int n0 = blockIdx.y * BN, nb0 = n0 + (int)rank * BNH;       // this CTA's half of the B rows
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // [0..2] full (leader only, count 2: one arrival per CTA); [3..5] free; [6] chunk done; [7..9] this CTA's B half has landed
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 128);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(smem_u32(&bars[i]), 2); mbar_init(smem_u32(&bars[3 + i]), 1); mbar_init(smem_u32(&bars[7 + i]), 1); }
    mbar_init(smem_u32(&bars[6]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // staging map (dense, k contiguous): thread -> 16-byte chunk kc = lane & 7 of rows (warp + 16 i) * 4 + (lane >> 3), i = 0, 1
  const int kc = lane & 7, rr = lane >> 3;
  float4 va[2];
  const int nkb = K / BK;
  const uint8_t* b_tiles = Bplanes + (int64_t)(nb0 / 128) * nkb * BHALF_BYTES;
  auto fetch = [&](int kb) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = (warp + i * 16) * 4 + rr;
      va[i] = __ldg(reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * K + kb * BK + kc * 4));
    }
  };
  auto load_b = [&](int kb) {      // thread 0 only
    const int sb_ = kb % NS;
    bulk_load(smem_u32(smem + sb_ * STAGE + 2 * PLANE), b_tiles + (int64_t)kb * BHALF_BYTES, BHALF_BYTES, smem_u32(&bars[7 + sb_]));
  };
  if (threadIdx.x == 0) { load_b(0); if (nkb > 1) load_b(1); }
  auto put = [&](uint8_t* hi_p, uint8_t* lo_p, int r, const float4& v) {
    float4 h, l;
    split(v.x, h.x, l.x); split(v.y, h.y, l.y); split(v.z, h.z, l.z); split(v.w, h.w, l.w);
    const uint32_t off = kc * LBO + r * 16;
    *reinterpret_cast<float4*>(hi_p + off) = h;
    *reinterpret_cast<float4*>(lo_p + off) = l;
  };
  fetch(0);
  const int q = warp & 3, part = warp >> 2;        // accumulator: row q*32 + lane, columns part*64 .. +63
  float acc[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) acc[j] = 0.f;
  int n_drained = 0;
  for (int it = 0; it < nkb; ++it) {
    const int s = it % NS;
    if (it >= NS) mbar_wait(smem_u32(&bars[3 + s]), (uint32_t)(((it / NS) - 1) & 1));
    uint8_t* sb = smem + s * STAGE;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = (warp + i * 16) * 4 + rr;
      put(sb, sb + PLANE, r, va[i]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const bool chunk_first = (it % DRAIN) == 0;
    const bool chunk_last = ((it + 1) % DRAIN) == 0 || it == nkb - 1;
    if (threadIdx.x == 0) {
      mbar_wait(smem_u32(&bars[7 + s]), (uint32_t)((it / NS) & 1));   // this CTA's B half of stage s has landed
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive_cluster(smem_u32(&bars[s]), 0);                  // this CTA's half of stage s is in place
      if (rank == 0) {
        mbar_wait(smem_u32(&bars[s]), (uint32_t)((it / NS) & 1));  // both halves are
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_u32(sb), a_lo = a_hi + PLANE, b_hi = a_hi + 2 * PLANE, b_lo = a_hi + 3 * PLANE;
#pragma unroll
        for (int j = 0; j < BK / 8; ++j) {
          const uint64_t da_hi = make_desc(a_hi + 2 * j * LBO, LBO, 128), da_lo = make_desc(a_lo + 2 * j * LBO, LBO, 128);
          const uint64_t db_hi = make_desc(b_hi + 2 * j * LBO, LBO, 128), db_lo = make_desc(b_lo + 2 * j * LBO, LBO, 128);
          mma2_tf32(tmem_base, da_lo, db_hi, IDESC, (!chunk_first || j > 0) ? 1u : 0u);
          mma2_tf32(tmem_base, da_hi, db_lo, IDESC, 1u);
          mma2_tf32(tmem_base, da_hi, db_hi, IDESC, 1u);
        }
        commit2(smem_u32(&bars[3 + s]));
        if (chunk_last) commit2(smem_u32(&bars[6]));
      }
      if (it + 2 < nkb) {      // buffer (it + 2) % NS was last read by stage it - 1: refill it, two stages of MMAs ahead
        const int s2 = (it + 2) % NS;
        if (it >= 1) mbar_wait(smem_u32(&bars[3 + s2]), (uint32_t)(((it - 1) / NS) & 1));
        load_b(it + 2);
      }
    }
    if (it + 1 < nkb) fetch(it + 1);
    if (chunk_last) {
      mbar_wait(smem_u32(&bars[6]), (uint32_t)(n_drained & 1));
      ++n_drained;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * 64 + cc * 32), v);
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
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
}

inline size_t planes128_bytes(int N, int K) { return (size_t)(N / 128) * (K / BK) * BHALF_BYTES; }
inline cudaError_t presplit128_launch(const float* B, int N, int K, uint8_t* planes, cudaStream_t st) {
  presplit128_kernel<<<148 * 16, 256, 0, st>>>(B, N, K, planes);
  return cudaGetLastError();
}
inline cudaError_t pair_tma_launch(const float* A, const uint8_t* Bplanes, float* C, int M, int N, int K, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(pair_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid(2 * (M / 256), N / 256);
  pair_tma_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(A, Bplanes, C, M, N, K);
  return cudaGetLastError();
}

}  // namespace pairtma
}  // namespace mvb
