// tc_gemm.cuh -- split-TF32 (3xTF32) tensor-core contraction core for sm_100a.
//
//   C[M,N] = sum_k A[m,k] * B[n,k]        (fp32 in, fp32 out, ~fp32 accuracy)
//
// Both operands are "separable gathers": element (r,k) = base[roff[r] + koff[k]].
// That one form covers every contraction on the ridge hot path:
//   * the implicit lag-expanded design  Xt[(trial,t),(c,w)] = Xpad[trial,c,t+w]
//     (rows: roff = trial*C*Tp + t, cols: koff = c*Tp + w)  -- the lag expansion
//     (reference: mvpy/estimators/timedelayed.py:394-431) is never materialised (X^T X itself has its own kernel,
//     toeplitz_gram.cuh, which needs no gather tables at all),
//   * dense row-major matrices (eigenvector / coefficient operands),
//   * feature masks and training-trial subsets (just shorter offset tables).
//
// Hardware mapping: tcgen05.mma kind::tf32, M=128 x N=BN tile per CTA, fp32
// accumulator in TMEM.  Each fp32 operand value is split in registers into
// hi = top 19 bits (exact TF32) and lo = x - hi; three MMAs per K-step
// (hi*hi + hi*lo + lo*hi) recover ~2^-21 relative accuracy.  Operand tiles are
// staged by the CTA's threads straight into the UMMA canonical K-major
// no-swizzle layout (8x16B core matrices; LBO padded by 16 B so the stores are
// bank-conflict free), double buffered, with tcgen05.commit -> mbarrier
// releasing a buffer back to the loaders (tc_gemm_kernel).
// Products whose B operand is handed over as pre-split planes (TcGemmParams::b_planes) run on tc_gemm_ta_kernel instead:
// A is split in registers and written to tensor memory (tcgen05.st), the MMAs read it from there, and shared memory
// holds a three-stage ring of B plane tiles filled by TMA bulk copies.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

namespace mvb {

struct GatherOperand {
  const float* base;      // device pointer
  const int64_t* roff;    // [rows] element offsets of the tile rows (M or N index)
  const int64_t* koff;    // [K]    element offsets along the contraction index
  int rows;               // number of valid rows
  int lanes_along_k;      // 0: consecutive rows are contiguous in memory; 1: consecutive k are;
                          // 2: as 1 and every 4-aligned group koff[4q..4q+3] is contiguous and 16-byte
                          //    aligned for every row (dense row-major, ld % 4 == 0, K % 4 == 0): 128-bit loads
};

struct TcGemmParams {
  GatherOperand A, B;
  int M, N, K;
  float* C;               // row-major [M][ldc] (+ z * c_split_stride for split-K partials)
  int64_t ldc;
  int64_t c_split_stride;
  int symmetric;          // 1: A==B; tiles strictly above the diagonal are skipped and mirrored
  int dbg;                // profiling only: 1 = skip operand staging, 2 = skip MMA issue (results invalid)
  // ---- batching (independent problems in one launch; blockIdx.z = batch * splits + split) ----
  int batch;                          // >= 1
  int64_t a_roff_batch, b_roff_batch; // per-batch stride (in entries) into the roff tables, 0 = shared
  int64_t a_koff_batch, b_koff_batch; // same for the koff tables
  int64_t c_batch_stride;             // per-batch offset of C (elements), used when c_roff == nullptr
  const int64_t* c_roff;              // optional [batch][M] element offsets of the output rows
  const unsigned int* skip_flag;      // optional: the whole launch is a no-op when *skip_flag != 0
  int raster_m_fast;                  // 1: consecutive CTAs walk the M tiles of one N tile (B tile stays hot in L2 while a
                                      //    large dense B streams once); 0: consecutive CTAs walk the N tiles of one M tile
  // ---- output options ----
  int c_block_cols;                   // > 0: C is stored in column blocks, [N / c_block_cols][M][c_block_cols]
  int64_t c_block_stride;             //      elements between consecutive column blocks (>= M * c_block_cols)
  float* rowsq;                       // optional [N tiles][TC_ROWSQ_PARTS][batch][M]: += sum over this thread's columns of C[m][n]^2
                                      // (split-K not allowed); C may be nullptr when only the row squares are wanted
  int n_tiles_per_cta;                // > 1 (tc_gemm_kernel, k_splits == 1, not symmetric): a CTA walks this many consecutive N tiles
  int rowsq_assign;                   // 1: rowsq entries are written (=) instead of accumulated (+=): every entry is produced once
  int b_tri_period;                   // > 0: B is a stack of lower-triangular blocks (B[n][k] = 0 for k > n mod period*BN), `period`
                                      //      N tiles each: tile t contracts only k < ((t mod period) + 1) * BN  (batch kernel only)
  // ---- pre-split B operand (256-wide tiles only, batch == 1) ----
  const uint8_t* b_planes;            // optional: B already split into TF32 hi / lo planes, one shared-memory image per
  int b_planes_nkb;                   //   (256-row tile, 32-k stage), [n_tile][b_planes_nkb][2 * plane bytes] (tc_presplit_launch);
                                      //   the stage's B half is then one TMA bulk copy instead of LDG -> split -> STS by the threads
  int b_stream_once;                  // 1 (TMEM-A kernel): every B tile is read once by one CTA (single M tile) while the small A operand is
                                      //    re-read by every CTA: B's bulk copies carry an L2 evict-first policy and A's loads evict-last, so
                                      //    that the B stream does not push A out of L2
                                      // 2: the other way round (raster_m_fast products: a B tile is re-read by all M tiles of a wave while the
                                      //    dense A operand streams): B evict-last, A evict-first
  int b_planes_raw;                   // 1: b_planes holds ONE plane per (tile, stage) -- the raw fp32 values in the same layout (4 B per
                                      //    entry instead of 8): the tensor core reads their top 19 bits as the hi part, the lo plane is
                                      //    produced in shared memory by the CTA's threads.  For products that stream B once from HBM
                                      //    (the skinny block-Lanczos products), which are bound by the plane bytes.
};

inline TcGemmParams tc_gemm_defaults() {
  TcGemmParams p{};
  p.batch = 1;
  static const int dbg = [] { const char* e = getenv("MVB_TC_DBG"); return e ? atoi(e) : 0; }();    // developer timing experiments only
  p.dbg = dbg;
  return p;
}

constexpr int TC_ROWSQ_PARTS = 4;      // = tc::NPART (column parts of the accumulator tile)

namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;          // fp32 elements per stage along K
constexpr int KC = BK / 4;      // 16-byte chunks per stage row
constexpr int NTHREADS = 512;      // 16 warps: the staging loop is latency- and issue-bound, so two CTAs' worth of warps
constexpr int NWARPS = NTHREADS / 32;
constexpr int NPART = NTHREADS / 128;   // column parts of the accumulator tile (one per group of four warps)
static_assert(NPART == TC_ROWSQ_PARTS, "rowsq layout");
// The tensor core adds into its fp32 accumulator with truncation, which biases long sums (measured:
// error grows linearly with K).  Every DRAIN stages the TMEM tile is therefore folded into fp32
// registers with round-to-nearest adds and the next MMA restarts the TMEM accumulator from zero.
constexpr int DRAIN = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
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
  printf("mvb tc_gemm: mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y,
         blockIdx.z, threadIdx.x);
  __trap();
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (layout_type 0), version 1.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

// One TMA bulk copy global -> shared (16-byte aligned, size % 16 == 0), completion counted on `bar` (bytes).
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]),
        "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]),
        "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]),
        "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// fp32 -> (hi, lo) with hi, lo both exactly representable in TF32 (round-to-nearest on the
// 13 dropped mantissa bits), so hi*hi + hi*lo + lo*hi carries ~2^-22 relative error per product
// with no truncation bias.
__device__ __forceinline__ float rn_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
// lo = x - hi is exact in fp32 and carries <= 13 significant bits; the tensor core reads its top 11 (truncation).  Because
// hi is rounded to nearest, lo is symmetric around zero, so that truncation is unbiased and <= 2^-21 |x|: rounding lo
// explicitly (two more integer ops per element in the staging loop, which is issue-bound) buys nothing measurable.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = rn_tf32(x);
  lo = x - hi;
}

template <int BN>
struct Cfg {
  static constexpr int A_LBO = BM * 16 + 16;
  static constexpr int B_LBO = BN * 16 + 16;
  static constexpr int A_PLANE = KC * A_LBO;
  static constexpr int B_PLANE = KC * B_LBO;
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr int BAR_OFF = 2 * STAGE;
  static constexpr int ROFF_OFF = BAR_OFF + 64;
  static constexpr int SMEM_BYTES = ROFF_OFF + (BM + BN) * 8;
  // instruction descriptor: D=F32, A=B=TF32, K-major both, N=BN, M=128
  static constexpr uint32_t IDESC =
      (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

// Operand tile loader.  ROWS x BK fp32 values per stage, fetched into registers one stage ahead
// (so the global/L2 latency overlaps the previous stage's MMAs) and then split + stored into the
// hi / lo planes.  ALONG_K: a warp walks one tile row with its lanes on consecutive k (coalesced when
// k is the contiguous direction); otherwise lanes sit on 32 consecutive rows and each thread owns one
// 16-byte K chunk (coalesced when rows are contiguous, e.g. lags of one feature).
template <int ROWS, int MODE>
struct Loader {
  static constexpr int NV = ROWS * BK / NTHREADS;
  static constexpr int LBO = ROWS * 16 + 16;
  static constexpr int RG_PER = (ROWS / 32) / (NWARPS / 8);   // mode 0: 32-row groups per warp
  float v[NV];

  __device__ __forceinline__ void fetch(const GatherOperand& op, const int64_t* sroff, int row0, int k0, int K) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (MODE == 2) {
      // lane -> (row-in-group-of-4, 16-byte chunk); 8 lanes read one row's 128 contiguous bytes
      const int kc = lane & 7, rr = lane >> 3;
      const int kg = k0 + kc * 4;
      const bool kok = kg < K;
      const float* src = op.base + (kok ? op.koff[kg] : 0);
#pragma unroll
      for (int i = 0; i < NV / 4; ++i) {
        const int r = (warp + i * (NTHREADS / 32)) * 4 + rr;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kok && row0 + r < op.rows) t = __ldg(reinterpret_cast<const float4*>(src + sroff[r]));
        v[i * 4 + 0] = t.x; v[i * 4 + 1] = t.y; v[i * 4 + 2] = t.z; v[i * 4 + 3] = t.w;
      }
    } else if (MODE == 1) {
      const int kg = k0 + lane;
      const bool kok = kg < K;
      const float* src = op.base + (kok ? op.koff[kg] : 0);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int r = warp + i * (NTHREADS / 32);
        v[i] = (kok && row0 + r < op.rows) ? __ldg(src + sroff[r]) : 0.f;
      }
    } else {
      int64_t ko[4];
      bool kok[4];
      const int kch = warp & 7, rg0 = (warp >> 3) * RG_PER;      // 8 k-chunks x (NWARPS / 8) row-group slices
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kg = k0 + kch * 4 + j;
        kok[j] = kg < K;
        ko[j] = kok[j] ? op.koff[kg] : 0;
      }
      // the warp's four k are usually consecutive in memory as well (lags of one feature, samples of one trial):
      // one address per row, the loads use immediate offsets (warp-uniform branch)
      const bool kcontig = kok[3] && (ko[1] - ko[0]) == 1 && (ko[2] - ko[0]) == 2 && (ko[3] - ko[0]) == 3;
      if (kcontig) {
#pragma unroll
        for (int rg = 0; rg < RG_PER; ++rg) {
          const int r = (rg0 + rg) * 32 + lane;
          const bool rok = row0 + r < op.rows;
          const float* src = op.base + (sroff[r] + ko[0]);
#pragma unroll
          for (int j = 0; j < 4; ++j) v[rg * 4 + j] = rok ? __ldg(src + j) : 0.f;
        }
      } else {
#pragma unroll
        for (int rg = 0; rg < RG_PER; ++rg) {
          const int r = (rg0 + rg) * 32 + lane;
          const bool rok = row0 + r < op.rows;
          const float* src = op.base + sroff[r];
#pragma unroll
          for (int j = 0; j < 4; ++j) v[rg * 4 + j] = (rok && kok[j]) ? __ldg(src + ko[j]) : 0.f;
        }
      }
    }
  }

  __device__ __forceinline__ void store(uint8_t* hi_plane, uint8_t* lo_plane) const {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (MODE == 2) {
      const int kc = lane & 7, rr = lane >> 3;
#pragma unroll
      for (int i = 0; i < NV / 4; ++i) {
        const int r = (warp + i * (NTHREADS / 32)) * 4 + rr;
        float4 h, l;
        split_tf32(v[i * 4 + 0], h.x, l.x);
        split_tf32(v[i * 4 + 1], h.y, l.y);
        split_tf32(v[i * 4 + 2], h.z, l.z);
        split_tf32(v[i * 4 + 3], h.w, l.w);
        const uint32_t off = kc * LBO + r * 16;
        *reinterpret_cast<float4*>(hi_plane + off) = h;
        *reinterpret_cast<float4*>(lo_plane + off) = l;
      }
    } else if (MODE == 1) {
      const uint32_t lane_off = (lane >> 2) * LBO + (lane & 3) * 4;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int r = warp + i * (NTHREADS / 32);
        float hi, lo;
        split_tf32(v[i], hi, lo);
        const uint32_t off = lane_off + (r >> 3) * 128 + (r & 7) * 16;
        *reinterpret_cast<float*>(hi_plane + off) = hi;
        *reinterpret_cast<float*>(lo_plane + off) = lo;
      }
    } else {
      const int kch = warp & 7, rg0 = (warp >> 3) * RG_PER;
#pragma unroll
      for (int rg = 0; rg < RG_PER; ++rg) {
        const int r = (rg0 + rg) * 32 + lane;
        float4 h, l;
        split_tf32(v[rg * 4 + 0], h.x, l.x);
        split_tf32(v[rg * 4 + 1], h.y, l.y);
        split_tf32(v[rg * 4 + 2], h.z, l.z);
        split_tf32(v[rg * 4 + 3], h.w, l.w);
        const uint32_t off = kch * LBO + r * 16;
        *reinterpret_cast<float4*>(hi_plane + off) = h;
        *reinterpret_cast<float4*>(lo_plane + off) = l;
      }
    }
  }
};

// Epilogue shared by the contraction kernels: the thread's NACC register accumulators (row q * 32 + lane of the tile,
// columns col_half * NACC ..) -> C, with the split-K / batch / column-block / row-square / mirror options of TcGemmParams.
template <int NACC>
__device__ __forceinline__ void store_tile(const TcGemmParams& p, const float (&acc)[NACC], int m0, int n0, int BN, int q, int lane,
                                           int col_half, int batch_id, int split_id) {
    const int gm = m0 + q * 32 + lane;
    float* Cz = p.C + (int64_t)split_id * p.c_split_stride;
    const int64_t crow_off = (gm < p.M) ? (p.c_roff ? p.c_roff[(int64_t)batch_id * p.M + gm]
                                                     : (int64_t)batch_id * p.c_batch_stride + (int64_t)gm * p.ldc)
                                        : 0;
    if (gm < p.M && p.rowsq) {
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < NACC; ++j)
        if (n0 + col_half * NACC + j < p.N) sq += acc[j] * acc[j];
      float* dst = p.rowsq + (((int64_t)(n0 / BN) * TC_ROWSQ_PARTS + col_half) * p.batch + batch_id) * p.M + gm;
      if (p.rowsq_assign) *dst = sq; else *dst += sq;
    }
    if (gm < p.M && p.C != nullptr) {
#pragma unroll
      for (int cc = 0; cc < NACC / 32; ++cc) {
        const int c0 = col_half * NACC + cc * 32;
        if (n0 + c0 < p.N) {
          float* crow = Cz + crow_off + n0 + c0;
          if (p.c_block_cols > 0) {
            const int cg = n0 + c0;   // 32-aligned, so the group never straddles a column block (block width % 32 == 0)
            crow = Cz + (int64_t)(cg / p.c_block_cols) * p.c_block_stride + (int64_t)gm * p.c_block_cols + (cg % p.c_block_cols);
          }
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
}

template <int BN, int A_MODE, int B_MODE>
__global__ void __launch_bounds__(NTHREADS, 1) tc_gemm_kernel(const TcGemmParams p) {
  using C_ = Cfg<BN>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int m0 = (p.raster_m_fast ? blockIdx.x : blockIdx.y) * BM;
  // n_tiles_per_cta > 1: the CTA walks that many consecutive N tiles of its M tile with barriers / tensor memory set up once
  // (short-K batched products, whose CTAs would otherwise be mostly prologue and epilogue); plain, non-symmetric launches only
  const int tiles_per_cta = p.n_tiles_per_cta > 1 ? p.n_tiles_per_cta : 1;
  const int nt0 = (p.raster_m_fast ? blockIdx.y : blockIdx.x) * tiles_per_cta;
  int n0 = nt0 * BN;
  if (p.symmetric && n0 > m0 + BM - 1) return;  // strictly above the diagonal: mirrored by the lower tile
  if (p.skip_flag && *p.skip_flag) return;
  const int splits = gridDim.z / p.batch;
  const int batch_id = blockIdx.z / splits, split_id = blockIdx.z - batch_id * splits;
  GatherOperand opA = p.A, opB = p.B;
  opA.roff += batch_id * p.a_roff_batch; opA.koff += batch_id * p.a_koff_batch;
  opB.roff += batch_id * p.b_roff_batch; opB.koff += batch_id * p.b_koff_batch;

  uint8_t* stage_base[2] = {smem, smem + C_::STAGE};
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C_::BAR_OFF);  // [0],[1]: buffer free; [2]: all MMAs done; [3],[4]: B planes landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C_::BAR_OFF + 48);
  constexpr bool B_TMA = (B_MODE == 3);
  constexpr uint32_t B_BYTES = 2 * C_::B_PLANE;
  int64_t* sroffA = reinterpret_cast<int64_t*>(smem + C_::ROFF_OFF);
  int64_t* sroffB = sroffA + BM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), 1);
    mbar_init(smem_u32(&bars[3]), 1);
    mbar_init(smem_u32(&bars[4]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"((uint32_t)BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int r = threadIdx.x; r < BM + BN; r += NTHREADS) {
    if (r < BM) sroffA[r] = (m0 + r < opA.rows) ? opA.roff[m0 + r] : 0;
    else if (!B_TMA) sroffB[r - BM] = (n0 + r - BM < opB.rows) ? opB.roff[n0 + r - BM] : 0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  constexpr int NACC = BN / NPART;             // columns per thread: warps 4i .. 4i+3 own column part i
  const int q = warp & 3;                      // TMEM lane quarter this warp may access
  const int col_half = warp >> 2;              // column part index, 0 .. NPART-1
  uint32_t gs = 0;                             // stages issued so far by this CTA (buffer / barrier parity across tiles)
  int n_drained = 0;
  for (int tt = 0; tt < tiles_per_cta; ++tt) {
  if (tt > 0) {
    n0 = (nt0 + tt) * BN;
    if (n0 >= p.N) break;
    __syncthreads();                           // every thread is past its last read of the previous tile's row offsets
    for (int r = threadIdx.x; r < BN; r += NTHREADS)
      if (!B_TMA) sroffB[r] = (n0 + r < opB.rows) ? opB.roff[n0 + r] : 0;
    __syncthreads();
  }
  const uint8_t* b_tiles = B_TMA ? p.b_planes + (int64_t)(n0 / BN) * p.b_planes_nkb * (int64_t)B_BYTES : nullptr;
  // K range of this split
  const int k_eff = p.b_tri_period > 0 ? min(p.K, (((n0 / BN) % p.b_tri_period) + 1) * BN) : p.K;
  const int nkb_total = (k_eff + BK - 1) / BK;
  const int kb_per = (nkb_total + splits - 1) / splits;
  const int kb_begin = split_id * kb_per;
  const int kb_end = min(nkb_total, kb_begin + kb_per);
  const int nkb = max(0, kb_end - kb_begin);

  Loader<BM, A_MODE> la;
  Loader<BN, B_TMA ? 2 : B_MODE> lb;             // unused (and compiled away) when the TMA delivers B
  if (nkb > 0) {
    la.fetch(opA, sroffA, m0, kb_begin * BK, p.K);
    if (!B_TMA) lb.fetch(opB, sroffB, n0, kb_begin * BK, p.K);
    else if (threadIdx.x == 0)
      bulk_load(smem_u32(stage_base[gs & 1u] + 2 * C_::A_PLANE), b_tiles + (int64_t)kb_begin * B_BYTES, B_BYTES, smem_u32(&bars[3 + (gs & 1u)]));
  }
  float acc[NACC];
#pragma unroll
  for (int j = 0; j < NACC; ++j) acc[j] = 0.f;

  for (int it = 0; it < nkb; ++it) {
    const uint32_t g = gs + (uint32_t)it;      // == it for a CTA that owns one tile
    const int s = (int)(g & 1u);
    if (g >= 2) mbar_wait(smem_u32(&bars[s]), (uint32_t)(((g >> 1) - 1) & 1u));
    uint8_t* sb = stage_base[s];
    if (p.dbg != 1) {
      la.store(sb, sb + C_::A_PLANE);
      if (!B_TMA) lb.store(sb + 2 * C_::A_PLANE, sb + 2 * C_::A_PLANE + C_::B_PLANE);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const bool chunk_first = (it % DRAIN) == 0;
    const bool chunk_last = ((it + 1) % DRAIN) == 0 || it == nkb - 1;
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (B_TMA) mbar_wait(smem_u32(&bars[3 + s]), (uint32_t)((g >> 1) & 1u));      // this stage's B planes have landed
      if (p.dbg != 2) {
        const uint32_t a_hi = smem_u32(sb), a_lo = a_hi + C_::A_PLANE;
        const uint32_t b_hi = a_hi + 2 * C_::A_PLANE, b_lo = b_hi + C_::B_PLANE;
#pragma unroll
        for (int j = 0; j < BK / 8; ++j) {
          const uint64_t da_hi = make_desc(a_hi + 2 * j * C_::A_LBO, C_::A_LBO, 128);
          const uint64_t da_lo = make_desc(a_lo + 2 * j * C_::A_LBO, C_::A_LBO, 128);
          const uint64_t db_hi = make_desc(b_hi + 2 * j * C_::B_LBO, C_::B_LBO, 128);
          const uint64_t db_lo = make_desc(b_lo + 2 * j * C_::B_LBO, C_::B_LBO, 128);
          mma_tf32(tmem_base, da_lo, db_hi, C_::IDESC, (!chunk_first || j > 0) ? 1u : 0u);
          mma_tf32(tmem_base, da_hi, db_lo, C_::IDESC, 1u);
          mma_tf32(tmem_base, da_hi, db_hi, C_::IDESC, 1u);
        }
      }
      mma_commit(smem_u32(&bars[s]));
      if (chunk_last) mma_commit(smem_u32(&bars[2]));
      if (B_TMA && it + 1 < nkb) {
        // the other buffer was last read by stage it - 1; once those MMAs are done the next stage's planes can land in it
        // while this stage's MMAs run
        const int s1 = s ^ 1;
        if (g >= 1) mbar_wait(smem_u32(&bars[s1]), (uint32_t)(((g - 1) >> 1) & 1u));
        bulk_load(smem_u32(stage_base[s1] + 2 * C_::A_PLANE), b_tiles + (int64_t)(kb_begin + it + 1) * B_BYTES, B_BYTES,
                  smem_u32(&bars[3 + s1]));
      }
    }
    if (it + 1 < nkb && p.dbg != 1) {  // next stage's global loads fly while this stage's MMAs run
      la.fetch(opA, sroffA, m0, (kb_begin + it + 1) * BK, p.K);
      if (!B_TMA) lb.fetch(opB, sroffB, n0, (kb_begin + it + 1) * BK, p.K);
    }
    if (chunk_last) {
      mbar_wait(smem_u32(&bars[2]), (uint32_t)(n_drained & 1));
      ++n_drained;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (p.dbg != 2) {
#pragma unroll
        for (int cc = 0; cc < NACC / 32; ++cc) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(col_half * NACC + cc * 32), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[cc * 32 + j] += v[j];
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
  }

  gs += (uint32_t)nkb;
  // ---- epilogue: register accumulators -> global ------------------------------------------------
  store_tile<NACC>(p, acc, m0, n0, BN, q, lane, col_half, batch_id, split_id);
  }   // tiles of this CTA
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN)
                 : "memory");
  }
}

// ---- A operand in tensor memory, B as pre-split planes by TMA ("TA" kernel) ---------------------------------------------
// With both operands' hi / lo planes in shared memory a 32-k stage costs 96 KB of staging writes + 144 KB of MMA operand
// reads against 128 B/clk of shared-memory bandwidth: 1875 cycles, more than the stage's 1536 cycles of MMAs.  Here A never
// touches shared memory: every thread splits the 8 values it holds of one tile row and writes them to tensor memory
// (tcgen05.st); the MMAs take A from TMEM (the [a_tmem] operand form) and shared memory carries only B (one 64 KB bulk copy
// in, 96 KB of MMA reads per stage).  The shared memory this frees holds a third B stage, so a stage's copy is issued two
// stages of MMAs ahead.  TMEM columns: 0..255 accumulator, 256 + 64 s: stage s of A (hi at +0, lo at +32).
// 256-wide tiles, batch == 1, A gathers 0 and 2 (mode 1 would need a transposing load; it keeps the
// shared-memory kernel).  Measured (tools/experimental/tc_gemm_tmema_test): dense 8192^3 214 TFLOP/s against 141.
struct CfgTA {
  static constexpr int BN = 256, NS = 3;
  static constexpr int B_LBO = Cfg<256>::B_LBO, B_PLANE = Cfg<256>::B_PLANE;
  static constexpr uint32_t B_BYTES = 2 * B_PLANE;
  static constexpr int BAR_OFF = NS * (int)B_BYTES;         // [0..2] stage free; [3] chunk done; [4..6] B planes landed
  static constexpr int SMEM_BYTES = BAR_OFF + 128;
  static constexpr uint32_t COL_A = 256, TMEM_COLS = 512;
};

__device__ __forceinline__ void mma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  const uint32_t* u = reinterpret_cast<const uint32_t*>(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(u[0]),
               "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
               : "memory");
}

__device__ __forceinline__ uint64_t l2_policy(bool evict_first) {
  uint64_t pol;
  if (evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_load_hint(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar), "l"(pol)
               : "memory");
}
__device__ __forceinline__ float4 ldg_hint(const float4* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}

// exact remainder of x after truncation to TF32 (what the tensor core sees of a raw fp32 operand), itself rounded to TF32
__device__ __forceinline__ float lo_after_trunc(float x) {
  return rn_tf32(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u));
}

template <int A_MODE, bool RAW_B = false>
__global__ void __launch_bounds__(NTHREADS, 1) tc_gemm_ta_kernel(const TcGemmParams p) {
  using C_ = CfgTA;
  static_assert(A_MODE == 0 || A_MODE == 2, "A gather modes 0 and 2 only");
  constexpr uint32_t SRC_BYTES = RAW_B ? (uint32_t)C_::B_PLANE : C_::B_BYTES;      // bytes of one (tile, stage) in b_planes
  // B ring: three hi / lo pairs (copy issued two stages ahead), or -- raw planes -- FIVE single planes (copy issued four stages
  // ahead) plus two lo planes produced in place.  A bulk copy that comes from HBM takes ~7 000 cycles to land whatever its size
  // (measured: a stage of the skinny Lanczos products takes 3 500 cycles with 64 KB pairs and with 32 KB raw planes alike when two
  // copies are in flight), so what shortens those products is more copies in flight per SM, which the halved stage makes room for.
  constexpr int NSB = RAW_B ? 5 : C_::NS;
  constexpr int AHEAD = NSB - 1;
  constexpr uint32_t SLOT_BYTES = RAW_B ? (uint32_t)C_::B_PLANE : C_::B_BYTES;
  constexpr int BAR_OFF = RAW_B ? 7 * C_::B_PLANE : C_::BAR_OFF;        // bars [0..2] stage's MMAs done; [3] chunk done; [4..] landed
  extern __shared__ __align__(1024) uint8_t smem[];
  const int m0 = (p.raster_m_fast ? blockIdx.x : blockIdx.y) * BM;
  const int n0 = (p.raster_m_fast ? blockIdx.y : blockIdx.x) * C_::BN;
  if (p.symmetric && n0 > m0 + BM - 1) return;  // strictly above the diagonal: mirrored by the lower tile
  if (p.skip_flag && *p.skip_flag) return;
  const int splits = gridDim.z, split_id = blockIdx.z;
  const GatherOperand opA = p.A;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 96);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4 + NSB; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(C_::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const int nkb_total = (p.K + BK - 1) / BK;
  const int kb_per = (nkb_total + splits - 1) / splits;
  const int kb_begin = split_id * kb_per;
  const int kb_end = min(nkb_total, kb_begin + kb_per);
  const int nkb = max(0, kb_end - kb_begin);

  const uint8_t* b_tiles = p.b_planes + ((int64_t)(n0 / C_::BN) * p.b_planes_nkb + kb_begin) * (int64_t)SRC_BYTES;
  const int q = warp & 3, part = warp >> 2;   // TMEM lane quarter; k columns part * 8 .. + 8 of a stage, accumulator columns part * 64 ..
  const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
  const bool rok = m0 + q * 32 + lane < opA.rows;
  const float* arow = opA.base + (rok ? opA.roff[m0 + q * 32 + lane] : 0);
  float av[8];
  const uint64_t pol_a = l2_policy(p.b_stream_once == 2), pol_b = l2_policy(p.b_stream_once != 2);
  auto fetch = [&](int it) {      // this thread's 8 values of stage it: row q * 32 + lane, k = (kb_begin + it) * BK + part * 8 ..
    // dbg 3 (timing experiment, results invalid): every stage re-reads stage 0's A values (L1 hits): what the loop costs without A latency
    // dbg 4: B copies of every stage come from the first stage's planes (L2 hits): what it costs without the HBM stream
    const int k0 = (kb_begin + (p.dbg == 3 ? 0 : it)) * BK + part * 8;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int kg = k0 + g * 4;
      if (A_MODE == 2) {          // 4-aligned k groups are contiguous and 16-byte aligned; K % 4 == 0
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rok && kg < p.K) t = p.b_stream_once ? ldg_hint(reinterpret_cast<const float4*>(arow + opA.koff[kg]), pol_a)
                                                 : __ldg(reinterpret_cast<const float4*>(arow + opA.koff[kg]));
        av[g * 4 + 0] = t.x; av[g * 4 + 1] = t.y; av[g * 4 + 2] = t.z; av[g * 4 + 3] = t.w;
      } else {                    // lanes sit on consecutive rows (contiguous in memory): coalesced scalar loads
        int64_t ko[4];
        bool kok[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          kok[j] = kg + j < p.K;
          ko[j] = kok[j] ? opA.koff[kg + j] : 0;
        }
        const bool kcontig = kok[3] && (ko[1] - ko[0]) == 1 && (ko[2] - ko[0]) == 2 && (ko[3] - ko[0]) == 3;
        if (kcontig) {            // warp-uniform: one address, immediate offsets
          const float* src = arow + ko[0];
#pragma unroll
          for (int j = 0; j < 4; ++j) av[g * 4 + j] = rok ? __ldg(src + j) : 0.f;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) av[g * 4 + j] = (rok && kok[j]) ? __ldg(arow + ko[j]) : 0.f;
        }
      }
    }
  };
  auto load_b = [&](int it) {     // thread 0 only
    const int sb = it % NSB;
    const int64_t src_it = p.dbg == 4 ? 0 : it;
    if (p.b_stream_once) bulk_load_hint(smem_u32(smem + sb * SLOT_BYTES), b_tiles + src_it * SRC_BYTES, SRC_BYTES, smem_u32(&bars[4 + sb]), pol_b);
    else bulk_load(smem_u32(smem + sb * SLOT_BYTES), b_tiles + src_it * SRC_BYTES, SRC_BYTES, smem_u32(&bars[4 + sb]));
  };
  if (nkb > 0) {
    if (threadIdx.x == 0) {
#pragma unroll
      for (int d = 0; d < AHEAD; ++d)
        if (nkb > d) load_b(d);
    }
    fetch(0);
  }
  constexpr int NACC = C_::BN / NPART;
  float acc[NACC];
#pragma unroll
  for (int j = 0; j < NACC; ++j) acc[j] = 0.f;
  int n_drained = 0;

  for (int it = 0; it < nkb; ++it) {
    const int s = it % C_::NS;
    if (!RAW_B) {
      if (it >= C_::NS) mbar_wait(smem_u32(&bars[s]), (uint32_t)(((it / C_::NS) - 1) & 1));   // stage s (A in TMEM, B in smem) is free
    } else if (it >= 2) {
      // the lo plane of this stage goes where stage it - 2's was: its MMAs must be done (which implies stage it - 3's, i.e. the
      // A stage in TMEM is free as well)
      mbar_wait(smem_u32(&bars[(it - 2) % C_::NS]), (uint32_t)(((it - 2) / C_::NS) & 1));
    }
    {
      float hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split_tf32(av[j], hi[j], lo[j]);
      tmem_st8(lane_addr + C_::COL_A + s * 64 + part * 8, hi);
      tmem_st8(lane_addr + C_::COL_A + s * 64 + 32 + part * 8, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    if (RAW_B) {
      // the stage's raw plane has landed (requested four stages ago): every thread turns 4 of its 2048 16-byte groups into the lo
      // plane next to it.  The raw plane itself is the hi operand -- kind::tf32 ignores the low 13 mantissa bits.
      mbar_wait(smem_u32(&bars[4 + it % NSB]), (uint32_t)((it / NSB) & 1));
      const float4* raw = reinterpret_cast<const float4*>(smem + (it % NSB) * C_::B_PLANE);
      float4* lop = reinterpret_cast<float4*>(smem + (NSB + (it & 1)) * C_::B_PLANE);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = threadIdx.x + NTHREADS * i;                   // (k-chunk, row) of the 8 x 256 groups
        const int off = (e >> 8) * (C_::B_LBO / 16) + (e & 255);
        const float4 v = raw[off];
        lop[off] = make_float4(lo_after_trunc(v.x), lo_after_trunc(v.y), lo_after_trunc(v.z), lo_after_trunc(v.w));
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    const bool chunk_first = (it % DRAIN) == 0;
    const bool chunk_last = ((it + 1) % DRAIN) == 0 || it == nkb - 1;
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      mbar_wait(smem_u32(&bars[4 + it % NSB]), (uint32_t)((it / NSB) & 1));                  // this stage's B planes have landed
      const uint32_t b_hi = smem_u32(smem + (it % NSB) * SLOT_BYTES);
      const uint32_t b_lo = RAW_B ? smem_u32(smem + (NSB + (it & 1)) * C_::B_PLANE) : b_hi + C_::B_PLANE;
      const uint32_t a_hi = tmem_base + C_::COL_A + s * 64, a_lo = a_hi + 32;
#pragma unroll
      for (int j = 0; j < BK / 8; ++j) {
        const uint64_t db_hi = make_desc(b_hi + 2 * j * C_::B_LBO, C_::B_LBO, 128);
        const uint64_t db_lo = make_desc(b_lo + 2 * j * C_::B_LBO, C_::B_LBO, 128);
        mma_tf32_ta(tmem_base, a_lo + j * 8, db_hi, Cfg<256>::IDESC, (!chunk_first || j > 0) ? 1u : 0u);
        mma_tf32_ta(tmem_base, a_hi + j * 8, db_lo, Cfg<256>::IDESC, 1u);
        mma_tf32_ta(tmem_base, a_hi + j * 8, db_hi, Cfg<256>::IDESC, 1u);
      }
      mma_commit(smem_u32(&bars[s]));
      if (chunk_last) mma_commit(smem_u32(&bars[3]));
      if (it + AHEAD < nkb) {   // ring slot (it + AHEAD) % NSB = (it - 1) % NSB was last read by stage it - 1
        if (it >= 1) mbar_wait(smem_u32(&bars[(it - 1) % C_::NS]), (uint32_t)(((it - 1) / C_::NS) & 1));
        load_b(it + AHEAD);
      }
    }
    if (it + 1 < nkb) fetch(it + 1);   // next stage's global loads fly while this stage's MMAs run
    if (chunk_last) {
      mbar_wait(smem_u32(&bars[3]), (uint32_t)(n_drained & 1));
      ++n_drained;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int cc = 0; cc < NACC / 32; ++cc) {
        float v[32];
        tmem_ld32(lane_addr + (uint32_t)(part * NACC + cc * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[cc * 32 + j] += v[j];
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
  }

  store_tile<NACC>(p, acc, m0, n0, C_::BN, q, lane, part, 0, split_id);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C_::TMEM_COLS) : "memory");
  }
}

// B operand -> TF32 hi / lo planes in the kernel's shared-memory image, one (256-row tile, 32-k stage) at a time:
// planes[(tile * nkb_total + kb) * 2 * PLANE + plane * PLANE + kc * LBO + r * 16 + (k % 4) * 4], zero-filled outside
// the operand.  One thread per (tile row, 16-byte chunk); rows [row_begin, row_end) and stages [kb_begin, kb_end) only,
// so a growing operand (a Lanczos basis) is extended in place.
template <bool RAW>
__global__ void presplit_kernel(const GatherOperand op, int K, int nkb_total, int row_begin, int row_end, int kb_begin,
                                int kb_end, uint8_t* __restrict__ planes) {
  constexpr int BN = 256, LBO = BN * 16 + 16, PLANE = KC * LBO;
  const int64_t n_rows = row_end - row_begin, n_kb = kb_end - kb_begin;
  const int64_t total = n_rows * n_kb * KC;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int kc = (int)(e % KC);
    const int64_t t = e / KC;
    const int row = row_begin + (int)(t % n_rows);
    const int kb = kb_begin + (int)(t / n_rows);
    const int tile = row / BN, r = row % BN;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int k = kb * BK + kc * 4;
    if (row < op.rows && k < K) {
      const float* src = op.base + op.roff[row];
      if (op.lanes_along_k == 2 && k + 3 < K) v = __ldg(reinterpret_cast<const float4*>(src + op.koff[k]));
      else {
        v.x = __ldg(src + op.koff[k]);
        if (k + 1 < K) v.y = __ldg(src + op.koff[k + 1]);
        if (k + 2 < K) v.z = __ldg(src + op.koff[k + 2]);
        if (k + 3 < K) v.w = __ldg(src + op.koff[k + 3]);
      }
    }
    if (RAW) {                                    // one plane: the values themselves (TcGemmParams::b_planes_raw)
      uint8_t* dst = planes + ((int64_t)tile * nkb_total + kb) * PLANE + kc * LBO + r * 16;
      *reinterpret_cast<float4*>(dst) = v;
      if (r == 0) *reinterpret_cast<float4*>(dst + BN * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    float4 h, l;
    split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
    uint8_t* dst = planes + ((int64_t)tile * nkb_total + kb) * (2 * PLANE) + kc * LBO + r * 16;
    *reinterpret_cast<float4*>(dst) = h;
    *reinterpret_cast<float4*>(dst + PLANE) = l;
    if (r == 0) {                                 // the 16 bytes of padding that end every chunk
      *reinterpret_cast<float4*>(dst + BN * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(dst + PLANE + BN * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

template <int BN, int AM, int BMODE>
inline cudaError_t launch_variant(const TcGemmParams& p, int k_splits, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<BN, AM, BMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg<BN>::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  unsigned nt = (p.N + BN - 1) / BN;
  const unsigned mt = (p.M + BM - 1) / BM;
  if (p.n_tiles_per_cta > 1) {
    if (k_splits != 1 || p.symmetric) return cudaErrorInvalidValue;
    nt = (nt + p.n_tiles_per_cta - 1) / p.n_tiles_per_cta;
  }
  dim3 grid(p.raster_m_fast ? mt : nt, p.raster_m_fast ? nt : mt, k_splits * (p.batch > 0 ? p.batch : 1));
  tc_gemm_kernel<BN, AM, BMODE><<<grid, NTHREADS, Cfg<BN>::SMEM_BYTES, st>>>(p);
  return cudaGetLastError();
}

// MVB_TC_TMEMA=0 in the environment keeps the shared-memory A path for the pre-split-B contractions (A/B comparisons).
inline bool ta_enabled() {
  static const bool on = [] { const char* e = getenv("MVB_TC_TMEMA"); return !(e && e[0] == '0'); }();
  return on;
}

template <int AM>
inline cudaError_t launch_ta(const TcGemmParams& p, int k_splits, cudaStream_t st) {
  const unsigned nt = (p.N + CfgTA::BN - 1) / CfgTA::BN, mt = (p.M + BM - 1) / BM;
  dim3 grid(p.raster_m_fast ? mt : nt, p.raster_m_fast ? nt : mt, k_splits);
  if (p.b_planes_raw) {
    constexpr int SMEM_RAW = 7 * CfgTA::B_PLANE + 128;      // five raw planes + two lo planes + barriers
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_ta_kernel<AM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_RAW);
    if (e != cudaSuccess) return e;
    tc_gemm_ta_kernel<AM, true><<<grid, NTHREADS, SMEM_RAW, st>>>(p);
    return cudaGetLastError();
  }
  cudaError_t e = cudaFuncSetAttribute(tc_gemm_ta_kernel<AM>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgTA::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  tc_gemm_ta_kernel<AM><<<grid, NTHREADS, CfgTA::SMEM_BYTES, st>>>(p);
  return cudaGetLastError();
}

template <int BN, int AM>
inline cudaError_t launch_b(const TcGemmParams& p, int k_splits, cudaStream_t st) {
  if (p.b_planes) {
    if (BN != 256 || p.batch > 1) return cudaErrorInvalidValue;
    if (AM != 1 && (ta_enabled() || p.b_planes_raw)) return launch_ta<AM == 1 ? 0 : AM>(p, k_splits, st);
    if (p.b_planes_raw) return cudaErrorInvalidValue;  // single raw planes exist for the TMEM-A kernel only
    if (p.symmetric) return cudaErrorInvalidValue;      // symmetric products with pre-split planes: TMEM-A kernel only
    return launch_variant<256, AM, 3>(p, k_splits, st);
  }
  switch (p.B.lanes_along_k) {
    case 0: return launch_variant<BN, AM, 0>(p, k_splits, st);
    case 1: return launch_variant<BN, AM, 1>(p, k_splits, st);
    default: return launch_variant<BN, AM, 2>(p, k_splits, st);
  }
}

template <int BN>
inline cudaError_t launch_a(const TcGemmParams& p, int k_splits, cudaStream_t st) {
  switch (p.A.lanes_along_k) {
    case 0: return launch_b<BN, 0>(p, k_splits, st);
    case 1: return launch_b<BN, 1>(p, k_splits, st);
    default: return launch_b<BN, 2>(p, k_splits, st);
  }
}

}  // namespace tc

// Split-K factor for `tiles` output tiles of `nkb` BK-stages each on `sms` SMs (one CTA per SM): minimise
// waves x (stages per CTA + fixed per-CTA cost) + the HBM time of the partial sums (s writes + s reads + 1 write of
// out_elems floats; one stage ~ 1 us, HBM ~ 6e6 bytes/us) -- e.g. 51 skinny tiles split 3 ways are 153 CTAs = two
// waves for 3% more parallelism while 8 ways fill three waves, but a 0.66 GB Gram is never worth splitting.
inline int tc_choose_splits(int64_t tiles, int nkb, int max_splits, int64_t out_elems, int sms = 148) {
  int best = 1;
  double best_cost = 1e300;
  for (int s = 1; s <= max_splits; ++s) {
    if (s > 1 && s > nkb / tc::DRAIN) break;                 // keep at least one accumulation chunk per split
    const int64_t waves = (tiles * s + sms - 1) / sms;
    double cost = (double)waves * ((double)((nkb + s - 1) / s) + 8.0);
    if (s > 1) cost += (2.0 * s + 1.0) * (double)out_elems * 4.0 / 6.0e6 + 3.0;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return best;
}

// Pre-split planes of a B operand (see TcGemmParams::b_planes): bytes needed, and the (partial) fill.
inline size_t tc_planes_bytes(int rows, int K, bool raw = false) {
  return (size_t)((rows + 255) / 256) * ((K + tc::BK - 1) / tc::BK) * (raw ? 1 : 2) * tc::Cfg<256>::B_PLANE;
}
inline cudaError_t tc_presplit_launch(const GatherOperand& op, int K, uint8_t* planes, cudaStream_t st, int row_begin = 0,
                                      int row_end = -1, int kb_begin = 0, int kb_end = -1, bool raw = false) {
  const int nkb = (K + tc::BK - 1) / tc::BK;
  if (row_end < 0) row_end = ((op.rows + 255) / 256) * 256;      // whole tiles: rows past the operand are zero-filled
  if (kb_end < 0) kb_end = nkb;
  const int64_t total = (int64_t)(row_end - row_begin) * (kb_end - kb_begin) * tc::KC;
  if (total <= 0) return cudaSuccess;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  if (raw) tc::presplit_kernel<true><<<grid, 256, 0, st>>>(op, K, nkb, row_begin, row_end, kb_begin, kb_end, planes);
  else tc::presplit_kernel<false><<<grid, 256, 0, st>>>(op, K, nkb, row_begin, row_end, kb_begin, kb_end, planes);
  return cudaGetLastError();
}

// Launch helper.  bn = 256 for wide outputs, 128 otherwise; k_splits partial sums go to
// C + z * c_split_stride (caller reduces them).
inline cudaError_t tc_gemm_launch(const TcGemmParams& p, int k_splits, int bn, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0) return cudaSuccess;
  if (bn == 256) return tc::launch_a<256>(p, k_splits, st);
  return tc::launch_a<128>(p, k_splits, st);
}

}  // namespace mvb
