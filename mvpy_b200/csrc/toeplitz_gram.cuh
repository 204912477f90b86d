// toeplitz_gram.cuh -- X^T X of the implicit lag-expanded design straight from the padded stimulus: TMA + tcgen05 only.
//
//   G[(c,w),(c',w')] = sum_{trial,t} xp[trial,c,t+w] * xp[trial,c',t+w']        (mvpy/estimators/timedelayed.py:394-431
//                                                                                builds the operand; ridgecv.py forms X^T X)
//
// Both operands of this product are Toeplitz: element (row w, k = t) of a feature's block is x[t + w], a function of
// row + k.  The UMMA shared-memory descriptor (K-major, no swizzle) addresses an operand tile as
//     start + (row / 8) * SBO + (row % 8) * 16 B + (k / 4) * LBO + (k % 4) * 4 B,
// so with LBO = 16 B a row that starts 16 B after its neighbour IS a row shifted by four samples: eight rows
// w0, w0+4, ..., w0+28 of one feature are nothing but a (28 + K)-sample window of the series, and the tensor core reads
// the overlapping rows out of it by itself.  The four residues w mod 4 come from four copies of the stimulus shifted by
// 0..3 samples (so every window starts 16-byte aligned), the TF32 hi / lo split is elementwise and therefore done once on
// the stimulus: the "pack" is 2 planes x 4 shifts x the fold's padded training series = 8 x 16 MB at BASELINE configs[1], against
// the 1.98 GB lag-expanded design (4 GB as pre-split planes) that the previous Gram kernel streamed.
//
// Per CTA: one 128 x 256 tile of G in "virtual" row order -- the W lags of a feature are cut into blocks of 32
// (w = 32 b + 4 j + s), a block is 4 groups (s = 0..3) of 8 rows (j = 0..7), a group is one window.  W = 201 gives 7
// blocks per feature, the last one with 9 valid rows: 224 virtual rows per 201 real ones (the price of the 16-byte row
// pitch).  Per K chunk (<= 144 samples of one trial) the CTA receives 48 windows x 2 planes of <= 688 B by
// cp.async.bulk into a three-stage ring (96 threads issue one copy each), one thread issues the MMAs
// (lo*hi + hi*lo + hi*hi per 8 samples), the two 256-column TMEM accumulators alternate so that the fp32 fold of chunk
// c - 1 (tcgen05.ld + round-to-nearest adds, see tc_gemm.cuh: the tensor core accumulates with truncation) overlaps the
// MMAs of chunk c.  No thread ever touches an operand.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tc_gemm.cuh"

namespace mvb {
namespace tg {

constexpr int BM = 128, BN = 256;
constexpr int GA = BM / 8, GB = BN / 8;          // row groups (windows) per operand tile
constexpr int KMAX = 144;                        // longest K chunk (samples), multiple of 8
constexpr int WIN_F = 28 + KMAX;                 // window length in floats: rows start 0, 4, ..., 28 samples in
constexpr int WIN_B = WIN_F * 4;                 // 688 B, a multiple of 16 (= the descriptor's SBO)
constexpr int NS = 3;                            // ring depth
constexpr int STAGE_B = (GA + GB) * 2 * WIN_B;   // [A hi][A lo][B hi][B lo]
constexpr int BAR_OFF = NS * STAGE_B;
constexpr int SMEM_BYTES = BAR_OFF + 128;
constexpr int NCOPY = (GA + GB) * 2;             // bulk copies per chunk, one per thread 0..95
constexpr int NTHREADS = 512;
constexpr int PACK_SLACK = 64;                   // samples a window of the last (ragged) block may read past its series
static_assert(WIN_B % 16 == 0 && KMAX % 8 == 0, "window layout");

struct Params {
  const float* pack;          // [2 planes][4 shifts][n_trials][n_feat][Tpp]: plane_pl of series (trial, feature) shifted by s samples
  int64_t L;                  // n_trials * n_feat * Tpp, floats per (plane, shift)
  int n_trials, n_feat;
  int Tpp;                    // pitch of one series in the pack (multiple of 4, >= T + W - 1 + PACK_SLACK)
  int n_vblk, nblk_f, W, T;   // virtual 32-row blocks (n_feat * nblk_f), blocks per feature, lags, samples per trial
  int n_chunks, chunk_len;    // K chunks per trial: n_chunks - 1 of chunk_len samples, the last one takes the rest
  float* G;                   // [p][ldg], p = n_feat * W
  int64_t ldg;
};

// ---- the pack: TF32 planes of the four shifted copies of the selected (trial, feature) series of the padded stimulus ----
// pack[((pl * 4 + s) * n_series + series) * Tpp + j] = plane_pl(x[series][j + s]), zero past the series' Tp samples.
__global__ void pack_kernel(const float* __restrict__ xp, const int32_t* __restrict__ trial_idx, int n_trials, int64_t trial_stride,
                            const int32_t* __restrict__ feat_idx, int n_feat, int64_t feat_stride, int Tp, int Tpp,
                            float* __restrict__ pack) {
  const int64_t n_series = (int64_t)n_trials * n_feat, q4 = Tpp / 4;
  const int64_t L = n_series * Tpp, total = 4 * n_series * q4;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(e % q4) * 4;
    const int64_t t = e / q4;
    const int64_t series = t % n_series;
    const int s = (int)(t / n_series);
    const int i = (int)(series / n_feat), c = (int)(series - (int64_t)i * n_feat);
    const float* src = xp + (int64_t)(trial_idx ? trial_idx[i] : i) * trial_stride + (int64_t)(feat_idx ? feat_idx[c] : c) * feat_stride;
    float h[4], l[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int jj = j + s + u;
      tc::split_tf32(jj < Tp ? __ldg(src + jj) : 0.f, h[u], l[u]);
    }
    const int64_t dst = series * Tpp + j;
    *reinterpret_cast<float4*>(pack + (int64_t)s * L + dst) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(pack + (int64_t)(4 + s) * L + dst) = make_float4(l[0], l[1], l[2], l[3]);
  }
}

inline int pack_pitch(int Tp) { return (Tp + PACK_SLACK + 3) / 4 * 4; }
inline size_t pack_bytes(int n_trials, int n_feat, int Tp) { return (size_t)n_trials * n_feat * pack_pitch(Tp) * 8 * sizeof(float); }

// The pack (130 MB at configs[1]) is read ~450 times over the launch while 0.66 GB of G streams out through the same L2:
// windows are fetched with an evict-last policy and G is stored evict-first, so that the pack stays L2 resident
// (ncu without the hints: 1.82 GB of dram reads for this launch, 14 x the pack).
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_load_hint(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar), "l"(pol)
               : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void store_hint(float* p, float v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}

// Toeplitz operand descriptor: consecutive 16-byte K chunks of a row are 16 B apart (LBO), consecutive rows of a group too
// (fixed by the canonical layout), so a row is its neighbour shifted by four samples; groups are WIN_B apart (SBO).
__device__ __forceinline__ uint64_t toeplitz_desc(uint32_t saddr) { return tc::make_desc(saddr, 16, WIN_B); }

__global__ void __launch_bounds__(NTHREADS, 1) toeplitz_gram_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int mi = blockIdx.x, nj = blockIdx.y;
  if (8 * nj > 4 * mi + 3) return;                           // strictly above the (virtual) diagonal: mirrored by the lower tile
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);        // [0..2] stage full, [3..4] accumulator done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 64);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) tc::mbar_init(tc::smem_u32(&bars[i]), NCOPY);
    tc::mbar_init(tc::smem_u32(&bars[3]), 1);
    tc::mbar_init(tc::smem_u32(&bars[4]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // ---- this thread's copy (threads 0..95): one window of one plane per chunk ----
  int64_t copy_base = 0;        // flat offset inside a trial of the window's first sample (chunk offset added per chunk)
  uint32_t copy_dst = 0;        // byte offset inside a stage
  const float* copy_plane = p.pack;
  if (threadIdx.x < NCOPY) {
    const int pl = threadIdx.x / (GA + GB), grp = threadIdx.x % (GA + GB);
    int vb, s;
    if (grp < GA) { vb = 4 * mi + grp / 4; s = grp % 4; copy_dst = (uint32_t)((pl * GA + grp) * WIN_B); }
    else { const int gb = grp - GA; vb = 8 * nj + gb / 4; s = gb % 4; copy_dst = (uint32_t)((2 * GA + pl * GB + gb) * WIN_B); }
    if (vb >= p.n_vblk) vb = 0;                               // tile padding: any readable window, the rows are never stored
    copy_base = (int64_t)(vb / p.nblk_f) * p.Tpp + 32 * (vb % p.nblk_f);      // shift copy s starts s samples in: 16-byte aligned
    copy_plane = p.pack + (int64_t)(pl * 4 + s) * p.L;
  }
  const uint64_t pol = policy_evict_last(), pol_st = policy_evict_first();
  const int NC = p.n_trials * p.n_chunks;
  auto chunk_of = [&](int c, int& trial, int& k0, int& len) {
    trial = c / p.n_chunks;
    const int ci = c - trial * p.n_chunks;
    k0 = ci * p.chunk_len;
    len = (ci == p.n_chunks - 1) ? p.T - k0 : p.chunk_len;
  };
  auto issue_copy = [&](int c) {      // threads 0..95
    int trial, k0, len;
    chunk_of(c, trial, k0, len);
    const float* src = copy_plane + (int64_t)trial * p.n_feat * p.Tpp + copy_base + k0;
    const int st = c % NS;
    bulk_load_hint(tc::smem_u32(smem + st * STAGE_B + copy_dst), src, (uint32_t)((28 + len) * 4), tc::smem_u32(&bars[st]), pol);
  };
  if (threadIdx.x < NCOPY)
    for (int c = 0; c < NS && c < NC; ++c) issue_copy(c);

  const int q = warp & 3, part = warp >> 2;      // TMEM lane quarter = block of the M tile; accumulator columns part * 64 ..
  const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
  constexpr int NACC = BN / 4;
  float acc[NACC];
#pragma unroll
  for (int j = 0; j < NACC; ++j) acc[j] = 0.f;

  auto drain = [&](int c) {
    tc::mbar_wait(tc::smem_u32(&bars[3 + (c & 1)]), (uint32_t)((c >> 1) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int cc = 0; cc < NACC / 32; ++cc) {
      float v[32];
      tc::tmem_ld32(lane_addr + (uint32_t)((c & 1) * BN + part * NACC + cc * 32), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[cc * 32 + j] += v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  };

  for (int c = 0; c < NC; ++c) {
    if (threadIdx.x == 0) {
      int trial, k0, len;
      chunk_of(c, trial, k0, len);
      const int st = c % NS;
      tc::mbar_wait(tc::smem_u32(&bars[st]), (uint32_t)((c / NS) & 1));          // the chunk's 96 windows have landed
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sb = tc::smem_u32(smem + st * STAGE_B);
      const uint64_t a_hi = toeplitz_desc(sb), a_lo = toeplitz_desc(sb + GA * WIN_B);
      const uint64_t b_hi = toeplitz_desc(sb + 2 * GA * WIN_B), b_lo = toeplitz_desc(sb + (2 * GA + GB) * WIN_B);
      const uint32_t d = tmem_base + (uint32_t)((c & 1) * BN);
      const int nks = len >> 3;
      for (int ks = 0; ks < nks; ++ks) {          // 8 samples = 32 B = two 16-byte address units further into every window
        const uint64_t o = (uint64_t)(2 * ks);
        tc::mma_tf32(d, a_lo + o, b_hi + o, tc::Cfg<256>::IDESC, ks > 0 ? 1u : 0u);
        tc::mma_tf32(d, a_hi + o, b_lo + o, tc::Cfg<256>::IDESC, 1u);
        tc::mma_tf32(d, a_hi + o, b_hi + o, tc::Cfg<256>::IDESC, 1u);
      }
      tc::mma_commit(tc::smem_u32(&bars[3 + (c & 1)]));
    }
    if (c >= 1) {
      drain(c - 1);                                                       // overlaps the MMAs of chunk c
      if (c + 2 < NC && threadIdx.x < NCOPY) issue_copy(c + 2);           // stage (c - 1) % NS is free: its MMAs are done
    }
    __syncthreads();       // accumulator (c - 1) & 1 has been read by everyone before chunk c + 1 is issued into it
  }
  if (NC > 0) drain(NC - 1);

  // ---- epilogue: virtual (block, s, j) order -> rows / columns of G, mirror for the tiles that were skipped ----
  {
    const int vb_r = 4 * mi + q;
    const int s_r = lane >> 3, j_r = lane & 7;
    const int c_r = vb_r / p.nblk_f, b_r = vb_r - c_r * p.nblk_f;
    const int w_r = 32 * b_r + 4 * j_r + s_r;
    const bool row_ok = vb_r < p.n_vblk && w_r < p.W;
    const int64_t r = (int64_t)c_r * p.W + w_r;
#pragma unroll
    for (int bl = 0; bl < 2; ++bl) {
      const int vb_c = 8 * nj + 2 * part + bl;
      const int c_c = vb_c / p.nblk_f, b_c = vb_c - c_c * p.nblk_f;
      const bool blk_ok = row_ok && vb_c < p.n_vblk;
      const bool mirror = 8 * (vb_r / 8) > 4 * (vb_c / 4) + 3;           // tile holding (col, row) was skipped
#pragma unroll
      for (int jn = 0; jn < 8; ++jn)            // address order: w = 32 b + 4 j + s
#pragma unroll
        for (int sn = 0; sn < 4; ++sn) {
          const int w_c = 32 * b_c + 4 * jn + sn;
          if (blk_ok && w_c < p.W) {
            const int64_t col = (int64_t)c_c * p.W + w_c;
            const float v = acc[bl * 32 + sn * 8 + jn];
            store_hint(p.G + r * p.ldg + col, v, pol_st);
            if (mirror) store_hint(p.G + col * p.ldg + r, v, pol_st);      // a warp's 32 rows are 32 consecutive columns here: one line
          }
        }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// shapes this kernel takes: whole MMA K steps per trial, and enough tiles to fill the machine without split-K
inline bool usable(int n_time, int n_lag, int n_feat, bool any_size = false) {
  if (n_time % 8 != 0 || n_time < 8 || n_lag < 2) return false;
  if (any_size) return true;
  const int64_t nvb = (int64_t)n_feat * ((n_lag + 31) / 32);
  const int64_t mt = (nvb + 3) / 4;
  return mt * (mt + 2) / 4 >= 2 * 148;          // ~ lower-triangle tile count
}

inline void chunk_plan(int T, int& n_chunks, int& chunk_len) {
  n_chunks = (T + KMAX - 1) / KMAX;
  chunk_len = ((T + n_chunks - 1) / n_chunks + 7) / 8 * 8;
  while ((n_chunks - 1) * chunk_len >= T) --n_chunks;
}

inline cudaError_t launch(const Params& p, cudaStream_t st) {
  cudaError_t attr = cudaFuncSetAttribute(toeplitz_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (attr != cudaSuccess) return attr;
  dim3 grid((p.n_vblk + 3) / 4, (p.n_vblk + 7) / 8);
  toeplitz_gram_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(p);
  return cudaGetLastError();
}

}  // namespace tg
}  // namespace mvb
