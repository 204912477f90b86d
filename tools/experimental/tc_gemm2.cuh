// tc_gemm2.cuh -- warp-specialised split-TF32 (3xTF32) tensor-core contraction core for sm_100a.
//
//   C[M,N] = sum_k A[m,k] * B[n,k]        (fp32 in, fp32 out, ~fp32 accuracy)
//
// Same "separable gather" operands as tc_gemm.cuh (element (r,k) = x[roff[r] + koff[k]]), but every operand
// arrives PRE-SPLIT into two TF32-exact planes (hi = rn_tf32(x), lo = rn_tf32(x - hi); split_planes_kernel below),
// so staging a tile needs no arithmetic: loader warps copy 16-byte chunks global -> shared with cp.async, four
// stages deep, straight into the UMMA K-major 64-byte-swizzle layout.  Roles (288 threads, one CTA per SM):
//   warps 0-7   stage operands: cp.async (16 B when the chunk is contiguous and aligned, 4 x 4 B otherwise), commit
//               groups, fence.proxy.async, arrive on the stage's "full" mbarrier two stages behind the newest copies.
//               Every 128 k they also fold the finished TMEM tile into fp32 registers with round-to-nearest adds
//               (the tensor core's own accumulation truncates; measured error grows linearly with K otherwise).
//               TMEM holds two accumulator tiles and the fold of chunk i is scheduled where its MMAs are known to
//               be complete, so it overlaps the MMAs of chunk i+1 and the copies in flight.  Last: the epilogue.
//   warp 8      one lane issues tcgen05.mma kind::tf32 (lo*hi + hi*lo + hi*hi per 8-wide k-step); tcgen05.commit
//               releases the stage ("empty") and publishes finished chunks ("tmem full").
// Unaligned operands (the implicit lag-expanded design, whose rows start at every sample) keep 16-byte copies by
// storing four shifted copies of the small padded stimulus: copy s holds x[i + s], so element offset e is read
// from copy e & 3 at the aligned offset e - (e & 3).
// Chunks run along k (K-major tiles): operands whose k is not the contiguous direction still work through the
// 4-byte path, but callers keep transposed copies of such matrices instead.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace mvb {

struct PlaneOperand {
  const float* hi;        // TF32-exact high plane (device)
  const float* lo;        // TF32-exact low plane
  int64_t copy_stride;    // > 0: each plane holds 4 shifted copies, copy s at + s * copy_stride; 0: single copy
  const int64_t* roff;    // [rows] element offsets of the tile rows (M or N index)
  const int64_t* koff;    // [K]    element offsets along the contraction index
  int rows;               // number of valid rows
};

struct TcGemm2Params {
  PlaneOperand A, B;
  int M, N, K;
  float* C;               // row-major [M][ldc] (+ z * c_split_stride for split-K partials)
  int64_t ldc;
  int64_t c_split_stride;
  int symmetric;          // 1: A==B; tiles strictly above the diagonal are skipped and mirrored
  int batch;                          // >= 1 independent problems, blockIdx.z = batch * splits + split
  int64_t a_roff_batch, b_roff_batch; // per-batch stride (entries) into the roff tables, 0 = shared
  int64_t a_koff_batch, b_koff_batch;
  int64_t c_batch_stride;
  const int64_t* c_roff;              // optional [batch][M] element offsets of the output rows
  const unsigned int* skip_flag;      // optional: the launch is a no-op when *skip_flag != 0
};

inline TcGemm2Params tc_gemm2_defaults() {
  TcGemm2Params p{};
  p.batch = 1;
  return p;
}

// x[n] -> hi/lo planes, `copies` (1 or 4) shifted copies each: plane[s * copy_stride + i] = split(x[i + s])
__global__ void split_planes_kernel(const float* __restrict__ x, int64_t n, int copies, int64_t copy_stride,
                                    float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t total = (int64_t)copies * copy_stride;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = e / copy_stride, i = e - s * copy_stride;
    const float v = (i + s < n) ? x[i + s] : 0.f;
    const float h = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
    const float r = v - h;
    hi[e] = h;
    lo[e] = __uint_as_float((__float_as_uint(r) + 0x1000u) & 0xFFFFE000u);
  }
}

namespace tc2 {

constexpr int BM = 128;
constexpr int BK = 16;            // fp32 elements per stage along K (two 8-wide MMA k-steps)
constexpr int NS = 4;             // shared-memory stages
constexpr int LAG = 2;            // a stage is published this many stages behind the newest copies in flight
constexpr int DRAIN_KB = 8;       // stages per accumulation chunk (128 k)
constexpr int N_WORK_WARPS = 8;
constexpr int N_WORK_THREADS = N_WORK_WARPS * 32;
constexpr int NTHREADS = (N_WORK_WARPS + 1) * 32;   // 288

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a lost arrival traps (reported as a launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  printf("mvb tc_gemm2: mbarrier timeout (block %d,%d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y, blockIdx.z,
         threadIdx.x, bar, parity);
  __trap();
}

// UMMA shared-memory descriptor, K-major, version 1: sbo = distance between 8-row groups; lbo only matters without swizzle.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 0) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7) << 61;     // 0 none, 2 128B, 4 64B, 6 32B swizzle
  return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes, bool cache_l1) {
  if (cache_l1) asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
  else asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

template <int BN>
struct Cfg {
  static constexpr int A_PLANE = BM * BK * 4;          // 8 KB
  static constexpr int B_PLANE = BN * BK * 4;
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr int BAR_OFF = NS * STAGE;           // full[NS], empty[NS], tfull[2], tempty[2], tmem slot
  static constexpr int ROFF_OFF = BAR_OFF + 128;
  static constexpr int SMEM_BYTES = ROFF_OFF + (BM + BN) * 8;
  static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

// k offsets of one thread's 16-byte chunk (4 consecutive k) for one stage: first offset + "is contiguous"; the
// 4-byte path re-reads the individual offsets
struct KChunk {
  int64_t ko0;
  int k, kvalid;
  bool contig;
};
__device__ __forceinline__ KChunk fetch_kchunk(const PlaneOperand& op, int k, int K) {
  KChunk c;
  c.k = k;
  c.kvalid = min(4, max(0, K - k));
  c.ko0 = c.kvalid > 0 ? __ldg(op.koff + k) : 0;
  c.contig = false;
  if (c.kvalid == 4) c.contig = (__ldg(op.koff + k + 3) - c.ko0) == 3 && (__ldg(op.koff + k + 1) - c.ko0) == 1;
  return c;
}

// One operand's share of one stage (ROWS x BK elements, both planes), issued by the 256 worker threads.
//   plane layout (UMMA K-major, 64-byte swizzle): a tile row is 64 contiguous bytes (16 k); its 16-byte chunk c sits at
//   byte r * 64 + ((c ^ ((r >> 1) & 3)) * 16)  -- Swizzle<2,4,3> on the byte address.
// Four lanes copy one row (64 contiguous bytes in global AND in shared memory), a warp instruction covers 8 rows =
// 512 contiguous shared bytes: every LDGSTS wavefront is a full 128-byte line (the unswizzled interleaved layout
// scattered each row's chunks and cost ~21 wavefronts per instruction, ncu profiles/r01).
template <int ROWS>
__device__ __forceinline__ void load_operand(const PlaneOperand& op, const int64_t* sroff, int row0, const KChunk& kc,
                                             uint32_t hi_plane, uint32_t lo_plane) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool copies = op.copy_stride > 0;
  const int c = lane & 3, r8 = lane >> 2;
#pragma unroll
  for (int gi = 0; gi < ROWS / 8 / N_WORK_WARPS; ++gi) {
    const int r = (gi * N_WORK_WARPS + warp) * 8 + r8;
    const bool rok = row0 + r < op.rows;
    const int64_t ro = sroff[r];
    const int64_t e = ro + kc.ko0;
    const uint32_t dst = (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
    const int s = (int)(e & 3);
    if (kc.contig && (copies || s == 0)) {
      const int64_t src = rok ? (copies ? (int64_t)s * op.copy_stride + (e - s) : e) : 0;
      const uint32_t nb = rok ? 16u : 0u;
      cp_async16(hi_plane + dst, op.hi + src, nb, copies);
      cp_async16(lo_plane + dst, op.lo + src, nb, copies);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ok = rok && i < kc.kvalid;
        const int64_t src = ok ? ro + __ldg(op.koff + kc.k + i) : 0;
        cp_async4(hi_plane + dst + 4 * i, op.hi + src, ok ? 4u : 0u);
        cp_async4(lo_plane + dst + 4 * i, op.lo + src, ok ? 4u : 0u);
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1) tc_gemm2_kernel(const TcGemm2Params p) {
  using C_ = Cfg<BN>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  if (p.symmetric && n0 > m0 + BM - 1) return;  // strictly above the diagonal: mirrored by the lower tile
  if (p.skip_flag && *p.skip_flag) return;
  const int splits = gridDim.z / p.batch;
  const int batch_id = blockIdx.z / splits, split_id = blockIdx.z - batch_id * splits;
  PlaneOperand opA = p.A, opB = p.B;
  opA.roff += batch_id * p.a_roff_batch; opA.koff += batch_id * p.a_koff_batch;
  opB.roff += batch_id * p.b_roff_batch; opB.koff += batch_id * p.b_koff_batch;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C_::BAR_OFF);
  uint64_t* full = bars;            // [NS]  count = worker threads
  uint64_t* empty = bars + NS;      // [NS]  count = 1 (tcgen05.commit)
  uint64_t* tfull = bars + 2 * NS;  // [2]   count = 1 (tcgen05.commit)
  uint64_t* tempty = tfull + 2;     // [2]   count = worker warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 4);
  int64_t* sroffA = reinterpret_cast<int64_t*>(smem + C_::ROFF_OFF);
  int64_t* sroffB = sroffA + BM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(smem_u32(&full[s]), N_WORK_THREADS); mbar_init(smem_u32(&empty[s]), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&tfull[b]), 1); mbar_init(smem_u32(&tempty[b]), N_WORK_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)(2 * BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int r = threadIdx.x; r < BM + BN; r += NTHREADS) {
    if (r < BM) sroffA[r] = (m0 + r < opA.rows) ? opA.roff[m0 + r] : 0;
    else sroffB[r - BM] = (n0 + r - BM < opB.rows) ? opB.roff[n0 + r - BM] : 0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // K range of this split
  const int nkb_total = (p.K + BK - 1) / BK;
  const int kb_per = (nkb_total + splits - 1) / splits;
  const int kb_begin = split_id * kb_per;
  const int nkb = max(0, min(nkb_total, kb_begin + kb_per) - kb_begin);
  const int nchunks = (nkb + DRAIN_KB - 1) / DRAIN_KB;

  if (warp < N_WORK_WARPS) {
    // ================= stage operands, fold finished chunks, epilogue =================
    constexpr int NACC = BN / 2;                 // columns per thread: warps 0-3 low half, 4-7 high half
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int col_half = warp >> 2;
    float acc[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) acc[j] = 0.f;
    const uint32_t smem0 = smem_u32(smem);
    const int kc_lane = (lane & 3) * 4;
    int next_chunk = 0;

    auto fold_chunk = [&](int ch) {
      const int buf = ch & 1;
      mbar_wait(smem_u32(&tfull[buf]), (uint32_t)((ch >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + col_half * NACC);
#pragma unroll
      for (int cc = 0; cc < NACC / 32; ++cc) {
        float v0[16], v1[16];
        tmem_ld16(tbase + cc * 32, v0);
        tmem_ld16(tbase + cc * 32 + 16, v1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) { acc[cc * 32 + j] += v0[j]; acc[cc * 32 + 16 + j] += v1[j]; }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty[buf]));
    };

    KChunk ka, kb;
    if (nkb > 0) {
      ka = fetch_kchunk(opA, kb_begin * BK + kc_lane, p.K);
      kb = fetch_kchunk(opB, kb_begin * BK + kc_lane, p.K);
    }
    for (int it = 0; it < nkb + LAG; ++it) {
      if (it < nkb) {
        const int s = it % NS, u = it / NS;
        if (u > 0) mbar_wait(smem_u32(&empty[s]), (uint32_t)((u - 1) & 1));
        const KChunk ca = ka, cb = kb;
        if (it + 1 < nkb) {     // next stage's k offsets fly while this stage's copies are issued
          ka = fetch_kchunk(opA, (kb_begin + it + 1) * BK + kc_lane, p.K);
          kb = fetch_kchunk(opB, (kb_begin + it + 1) * BK + kc_lane, p.K);
        }
        const uint32_t sb = smem0 + s * C_::STAGE;
        load_operand<BM>(opA, sroffA, m0, ca, sb, sb + C_::A_PLANE);
        load_operand<BN>(opB, sroffB, n0, cb, sb + 2 * C_::A_PLANE, sb + 2 * C_::A_PLANE + C_::B_PLANE);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (it >= LAG) {
        asm volatile("cp.async.wait_group %0;" ::"n"(LAG) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(smem_u32(&full[(it - LAG) % NS]));
      }
      // having passed the empty-wait of stage `it`, the MMAs of stage it - NS are complete: a chunk whose last
      // stage is <= it - NS can be folded without waiting
      if (it < nkb && next_chunk < nchunks && (next_chunk + 1) * DRAIN_KB - 1 <= it - NS) { fold_chunk(next_chunk); ++next_chunk; }
    }
    for (; next_chunk < nchunks; ++next_chunk) fold_chunk(next_chunk);
    // ---- epilogue: register accumulators -> global ----
    const int gm = m0 + q * 32 + lane;
    float* Cz = p.C + (int64_t)split_id * p.c_split_stride;
    if (gm < p.M) {
      const int64_t crow_off = p.c_roff ? p.c_roff[(int64_t)batch_id * p.M + gm]
                                        : (int64_t)batch_id * p.c_batch_stride + (int64_t)gm * p.ldc;
#pragma unroll
      for (int cc = 0; cc < NACC / 32; ++cc) {
        const int c0 = col_half * NACC + cc * 32;
        if (n0 + c0 < p.N) {
          float* crow = Cz + crow_off + n0 + c0;
          const int nvalid = min(32, p.N - (n0 + c0));
          if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(crow + j) =
                  make_float4(acc[cc * 32 + j], acc[cc * 32 + j + 1], acc[cc * 32 + j + 2], acc[cc * 32 + j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid) crow[j] = acc[cc * 32 + j];
          }
          if (p.symmetric) {
            // mirror into tiles that were skipped (row-tile of gn lies strictly above col-tile of gm)
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int gn = n0 + c0 + j;
              if (j < nvalid && gn < p.M && gm < p.N) {
                const int mt0 = (gn / BM) * BM, nt0 = (gm / BN) * BN;
                if (nt0 > mt0 + BM - 1) Cz[(int64_t)gn * p.ldc + gm] = acc[cc * 32 + j];
              }
            }
          }
        }
      }
    }
  } else if (lane == 0) {
    // ================= MMA issuer =================
    const uint32_t smem0 = smem_u32(smem);
    constexpr uint32_t LBO = 16, SBO = 512, SWZ64 = 4;      // 8-row groups are 512 B apart; LBO is unused with swizzle
    for (int it = 0; it < nkb; ++it) {
      const int s = it % NS, u = it / NS;
      const int ch = it / DRAIN_KB, buf = ch & 1;
      const bool chunk_first = (it % DRAIN_KB) == 0;
      const bool chunk_last = ((it + 1) % DRAIN_KB) == 0 || it == nkb - 1;
      if (chunk_first && ch >= 2) mbar_wait(smem_u32(&tempty[buf]), (uint32_t)(((ch >> 1) - 1) & 1));
      mbar_wait(smem_u32(&full[s]), (uint32_t)(u & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem0 + s * C_::STAGE, a_lo = a_hi + C_::A_PLANE;
      const uint32_t b_hi = a_hi + 2 * C_::A_PLANE, b_lo = b_hi + C_::B_PLANE;
      const uint32_t td = tmem_base + (uint32_t)(buf * BN);
#pragma unroll
      for (int j = 0; j < BK / 8; ++j) {
        const uint64_t da_hi = make_desc(a_hi + 32 * j, LBO, SBO, SWZ64);   // k-step j = bytes [32 j, 32 j + 32) of each row
        const uint64_t da_lo = make_desc(a_lo + 32 * j, LBO, SBO, SWZ64);
        const uint64_t db_hi = make_desc(b_hi + 32 * j, LBO, SBO, SWZ64);
        const uint64_t db_lo = make_desc(b_lo + 32 * j, LBO, SBO, SWZ64);
        mma_tf32(td, da_lo, db_hi, C_::IDESC, (!chunk_first || j > 0) ? 1u : 0u);
        mma_tf32(td, da_hi, db_lo, C_::IDESC, 1u);
        mma_tf32(td, da_hi, db_hi, C_::IDESC, 1u);
      }
      mma_commit(smem_u32(&empty[s]));
      if (chunk_last) mma_commit(smem_u32(&tfull[buf]));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)) : "memory");
  }
}

template <int BN>
inline cudaError_t launch_bn(const TcGemm2Params& p, int k_splits, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, k_splits * (p.batch > 0 ? p.batch : 1));
  tc_gemm2_kernel<BN><<<grid, NTHREADS, Cfg<BN>::SMEM_BYTES, st>>>(p);
  return cudaGetLastError();
}

}  // namespace tc2

// bn = 256 for wide outputs, 128 otherwise; k_splits partial sums go to C + z * c_split_stride (caller reduces them).
inline cudaError_t tc_gemm2_launch(const TcGemm2Params& p, int k_splits, int bn, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0) return cudaSuccess;
  if (bn == 256) return tc2::launch_bn<256>(p, k_splits, st);
  return tc2::launch_bn<128>(p, k_splits, st);
}

}  // namespace mvb
