// lanczos_band.cuh -- orthogonal reduction of the centred Gram to block-tridiagonal form on the tensor cores.
//
//   Gp = Qt^T T Qt,   Qt orthogonal (rows = basis vectors),  T block tridiagonal with 128 x 128 blocks.
//
// Why: the LOO sweep needs (G + alpha I)^-1 applied to every design row for the whole alpha grid
// (mvpy/estimators/ridgecv.py:230-251).  The reference gets that from a full SVD; any orthogonal Q with banded
// Q^T G Q serves equally well -- (G + aI)^-1 = Q (T + aI)^-1 Q^T, and block-tridiagonal systems are solved by a
// block Cholesky recurrence -- and costs ~6 p^3 tensor-core flops instead of an eigensolver.
//
// Method: block Lanczos with full re-orthogonalisation, block size 128.  Per step:
//   W = Q_j G;  H = W [Q_{j-1}; Q_j]^T  = T_{j,j-1}, T_jj;  W -= H [Q_{j-1}; Q_j]       (three-term recurrence)
//   Q' = L^-1 W, W W^T = L L^T (Cholesky QR; eigen-based SVQB when a pivot says the block is ill conditioned)
//   Q' -= (Q' Q^T) Q against ALL previous blocks (the full pass, on the normalised block)
//   Q_{j+1} = (Q' Q'^T)^-1/2 Q'  (4th-order series; SVQB fallback when |Q'Q'^T - I| >= 0.03)
// MVB_LB_PASS1_LOCAL=0 / MVB_LB_REORTH2=1 restore the two extra full passes on W (measured: same accuracy,
// orth 3e-6, reconstruction 9e-7 at p = 603 ... 12 864; 272 ms instead of 166 ms at p = 12 864).
// Because every new block is orthonormal, orthogonal to all previous blocks and spans the residual W, T is block
// tridiagonal for any such choice -- rank-deficient residuals simply get noise directions, which is still a valid
// orthogonal completion.  Every step is a launch of the split-TF32 tcgen05 contraction core (tc_gemm.cuh) or a
// small helper kernel; nothing synchronises with the host.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <algorithm>
#include "tc_gemm.cuh"
#include "jacobi_block.cuh"   // small_eig_kernel (128 x 128 symmetric eigensolver in shared memory)

namespace mvb {
namespace lb {

constexpr int NB = 128;   // block size (= one MMA M tile, = small_eig size)
static_assert(NB == bj::PR, "block size must match the shared-memory eigensolver");

struct Plan {
  int p, P, nblk;
  size_t bytes;
  size_t off_Gp, off_Qc, off_Wt, off_Wt2, off_H, off_S, off_J, off_M, off_lam, off_rowP, off_row128, off_iota, off_part, off_ctl;
  size_t part_elems;
  // pre-split TF32 planes (tc_gemm.cuh, TcGemmParams::b_planes) of the three dense B operands of the big products:
  // Gp (filled once), Qt and Q^T (extended by one block per step); 0 bytes for small problems
  size_t off_plG, off_plQt, off_plQc, planes_bytes;
  bool planes_raw;         // planes hold the raw fp32 values (4 B per entry; TcGemmParams::b_planes_raw) instead of hi / lo pairs
};

inline Plan make_plan(int p) {
  Plan pl{};
  pl.p = p;
  pl.nblk = (p + NB - 1) / NB;
  if (pl.nblk < 1) pl.nblk = 1;
  pl.P = pl.nblk * NB;
  const size_t P = pl.P;
  size_t off = 0;
  auto take = [&](size_t bytes) { off = (off + 255) / 256 * 256; size_t o = off; off += bytes; return o; };
  pl.off_Gp = take(P * P * 4);
  pl.off_Qc = take(P * P * 4);          // Q^T, kept in step with Qt so that products over Q's rows read k-contiguous operands
  pl.off_Wt = take(NB * P * 4);
  pl.off_Wt2 = take(NB * P * 4);
  pl.off_H = take(NB * P * 4);
  pl.off_S = take(NB * NB * 4);
  pl.off_J = take(NB * NB * 4);
  pl.off_M = take(NB * NB * 4);
  pl.off_lam = take((NB + 8) * 4);
  pl.off_rowP = take(P * 8);
  pl.off_row128 = take(NB * 8);
  pl.off_iota = take(P * 8);
  pl.part_elems = (size_t)64 * NB * (P > 4096 ? 4096 : P);     // split-K partials (see gemm())
  if (pl.part_elems < (size_t)8 * NB * P) pl.part_elems = (size_t)8 * NB * P;
  pl.off_part = take(pl.part_elems * 4);
  pl.off_ctl = take(sizeof(bj::Ctl));
  static const bool planes_on = [] { const char* e = getenv("MVB_LB_PLANES"); return !e || atoi(e) != 0; }();
  // the big products of a step stream their B operand from HBM once against a single 128-row tile, i.e. they are bound by the
  // plane bytes: single raw planes halve them (MVB_LB_RAW_PLANES=0: hi / lo pairs as in r01)
  static const bool raw_on = [] { const char* e = getenv("MVB_LB_RAW_PLANES"); return e && atoi(e) != 0; }();
  pl.planes_raw = raw_on;
  pl.planes_bytes = (planes_on && pl.P >= 2048) ? tc_planes_bytes(pl.P, pl.P, raw_on) : 0;
  pl.off_plG = take(pl.planes_bytes);
  pl.off_plQt = take(pl.planes_bytes);
  pl.off_plQc = take(pl.planes_bytes);
  pl.bytes = off + 256;
  return pl;
}

// Gp = [[G, 0], [0, diag(pad)]], pad entries distinct and of the scale of G's diagonal
__global__ void pad_gram_kernel(const float* __restrict__ G, int p, int P, float* Gp, const float* diag_scale) {
  const int64_t total = (int64_t)P * P;
  const float g0 = fabsf(diag_scale[0]) > 0.f ? fabsf(diag_scale[0]) : 1.f;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / P, c = e - r * P;
    float v = 0.f;
    if (r < p && c < p) v = G[r * p + c];
    else if (r == c) v = g0 * (1.f + 0.01f * (float)(r - p));
    Gp[e] = v;
  }
}

// mean of G's diagonal (single block)
__global__ void diag_mean_kernel(const float* __restrict__ G, int p, float* out) {
  __shared__ float red[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < p; i += blockDim.x) s += G[(int64_t)i * p + i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) out[0] = red[0] / (float)p;
}

__global__ void tables_kernel(int P, int64_t* rowP, int64_t* row128, int64_t* iota) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    rowP[i] = (int64_t)i * P; iota[i] = i;
    if (i < NB) row128[i] = (int64_t)i * NB;
  }
}

// deterministic pseudo-random start block (hash of the element index), uniform in (-1, 1)
__global__ void random_block_kernel(float* W, int64_t n, uint32_t seed) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    uint32_t x = (uint32_t)e * 2654435761u + seed;
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    W[e] = (float)(x >> 8) * (2.0f / 16777216.0f) - 1.0f;
  }
}

__global__ void sub_kernel(float* __restrict__ W, const float* __restrict__ W2, int64_t n, const unsigned int* skip = nullptr) {
  if (skip && *skip) return;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) W[e] -= W2[e];
}

// out[z-sum] : reduce split-K partials (fixed order)
// subtract: out -= sum of partials (the Gram-Schmidt update W -= H Q without a temporary)
__global__ void reduce_parts_kernel(const float* __restrict__ part, int splits, int64_t stride, int M, int N, float* out, int64_t ldo,
                                    const unsigned int* skip = nullptr, int subtract = 0) {
  if (skip && *skip) return;
  const int64_t total = (int64_t)M * N;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = e / N, n = e - m * N;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[z * stride + e];
    out[m * ldo + n] = subtract ? out[m * ldo + n] - s : s;
  }
}

// Mmat = diag(lam^-1/2) J  (SVQB scaling; lam sorted like the rows of J, floored far below fp32 noise).
// guard != nullptr: no-op when guard->done is set (the Newton-Schulz matrix already in Mmat stands); clears the flag.
__global__ void svqb_scale_kernel(const float* __restrict__ J, const float* __restrict__ lam, float* Mmat, bj::Ctl* guard,
                                  const unsigned int* skip = nullptr) {
  __shared__ float lmax;
  if (skip && *skip) return;
  if (guard) {
    const unsigned int skip = guard->done;
    __syncthreads();
    if (threadIdx.x == 0) guard->done = 0;
    if (skip) return;
  }
  if (threadIdx.x == 0) { float m = 0.f; for (int r = 0; r < NB; ++r) m = fmaxf(m, lam[r]); lmax = m; }
  __syncthreads();
  const float floor_ = fmaxf(lmax * 1e-12f, 1e-36f);
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) Mmat[e] = J[e] * rsqrtf(fmaxf(lam[e / NB], floor_));
}

// Fast path of the block orthonormalisation: Mmat = L^-1 with S = L L^T (Cholesky QR), all in shared memory.
// Valid when S is well conditioned: sets guard->done = 1 when every pivot exceeds pivot_tol * max diag(S), else 0 so
// that the eigen-based SVQB (robust for rank-deficient residual blocks) runs instead.  guard->rot[2] counts fallbacks.
constexpr int CLD = NB + 1;
constexpr int CHOL_ORTH_SMEM = 2 * NB * CLD * 4;
__global__ void __launch_bounds__(1024) chol_orth_kernel(const float* __restrict__ S, float* __restrict__ Mmat, float pivot_tol,
                                                         bj::Ctl* guard) {
  extern __shared__ float csm[];
  float* Cm = csm;               // S -> L (lower)
  float* Li = csm + NB * CLD;    // L^-1
  __shared__ float s_min, s_max;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int r = e / NB, c = e % NB;
    Cm[r * CLD + c] = 0.5f * (S[e] + S[c * NB + r]);
    Li[r * CLD + c] = 0.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f;
    for (int r = 0; r < NB; ++r) m = fmaxf(m, Cm[r * CLD + r]);
    s_max = m; s_min = m;
  }
  __syncthreads();
  const float floor_ = fmaxf(s_max * 1e-10f, 1e-36f);
  // left-looking Cholesky in 16-column panels: 3 block barriers per panel instead of 3 per column
  constexpr int PB = 16;
  for (int pb = 0; pb < NB; pb += PB) {
    // (1) panel -= L[:, 0:pb] L[pb:pb+16, 0:pb]^T   (lower part only)
    for (int e = threadIdx.x; e < (NB - pb) * PB; e += blockDim.x) {
      const int i = pb + e / PB, c = pb + e % PB;
      if (c <= i) {
        float acc = 0.f;
        for (int k = 0; k < pb; ++k) acc += Cm[i * CLD + k] * Cm[c * CLD + k];
        Cm[i * CLD + c] -= acc;
      }
    }
    __syncthreads();
    // (2) factor the 16 x 16 diagonal block: one warp, lane r owns row pb + r
    if (threadIdx.x < 32) {
      const int r = threadIdx.x;
      float pmin = s_min;
      for (int k = 0; k < PB; ++k) {
        const float piv = Cm[(pb + k) * CLD + pb + k];
        const float d = sqrtf(fmaxf(piv, floor_));
        pmin = fminf(pmin, piv);
        __syncwarp();
        if (r == k) Cm[(pb + k) * CLD + pb + k] = d;
        if (r > k && r < PB) Cm[(pb + r) * CLD + pb + k] /= d;
        __syncwarp();
        if (r > k && r < PB) {
          const float lrk = Cm[(pb + r) * CLD + pb + k];
          for (int c = k + 1; c <= r; ++c) Cm[(pb + r) * CLD + pb + c] -= lrk * Cm[(pb + c) * CLD + pb + k];
        }
        __syncwarp();
      }
      if (r == 0) s_min = pmin;
    }
    __syncthreads();
    // (3) rows below the panel: L[i, panel] = A[i, panel] L_pp^-T, one thread per row
    for (int i = pb + PB + threadIdx.x; i < NB; i += blockDim.x) {
      float x[PB];
#pragma unroll
      for (int c = 0; c < PB; ++c) {
        float v = Cm[i * CLD + pb + c];
#pragma unroll
        for (int k = 0; k < PB; ++k)
          if (k < c) v -= x[k] * Cm[(pb + c) * CLD + pb + k];
        x[c] = v / Cm[(pb + c) * CLD + pb + c];
      }
#pragma unroll
      for (int c = 0; c < PB; ++c) Cm[i * CLD + pb + c] = x[c];
    }
    __syncthreads();
  }
  // Li = L^-1 (lower) by 16 x 16 blocks:  M_ii = L_ii^-1 (all eight diagonal blocks at once, one thread per column),
  // then block row by block row  M_ij = -M_ii sum_{j <= k < i} L_ik M_kj  (two barriers per block row).
  constexpr int NBLK = NB / PB;
  if (threadIdx.x < NB) {
    const int blk = threadIdx.x / PB, c = threadIdx.x % PB, o = blk * PB;      // column c of diagonal block blk
    float x[PB];
#pragma unroll
    for (int i = 0; i < PB; ++i) {
      float v = (i == c) ? 1.f : 0.f;
#pragma unroll
      for (int k = 0; k < PB; ++k)
        if (k < i && k >= c) v -= Cm[(o + i) * CLD + o + k] * x[k];
      x[i] = (i < c) ? 0.f : v / Cm[(o + i) * CLD + o + i];
    }
#pragma unroll
    for (int i = 0; i < PB; ++i) Li[(o + i) * CLD + o + c] = x[i];
  }
  __syncthreads();
  float* Tm = Cm;      // the strictly-lower blocks of Cm are read (L_ik) and, per block row, reused for T_ij after the reads
  for (int bi = 1; bi < NBLK; ++bi) {
    // T_ij = sum_{k=j}^{bi-1} L_{bi,k} M_{k,j}  for all j < bi: bi * 256 elements
    float t_acc[2] = {0.f, 0.f};
    int n_mine = 0;
    for (int e = threadIdx.x; e < bi * PB * PB; e += blockDim.x, ++n_mine) {
      const int j = e / (PB * PB), r = (e / PB) % PB, c = e % PB;
      float acc = 0.f;
      for (int k = j * PB; k < bi * PB; ++k) acc += Cm[(bi * PB + r) * CLD + k] * Li[k * CLD + j * PB + c];
      t_acc[n_mine] = acc;
    }
    __syncthreads();
    n_mine = 0;
    for (int e = threadIdx.x; e < bi * PB * PB; e += blockDim.x, ++n_mine) {
      const int j = e / (PB * PB), r = (e / PB) % PB, c = e % PB;
      Tm[(bi * PB + r) * CLD + j * PB + c] = t_acc[n_mine];       // overwrite L_{bi,j} with T_{bi,j} (all reads are done)
    }
    __syncthreads();
    for (int e = threadIdx.x; e < bi * PB * PB; e += blockDim.x) {
      const int j = e / (PB * PB), r = (e / PB) % PB, c = e % PB;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < PB; ++k) acc += Li[(bi * PB + r) * CLD + bi * PB + k] * Tm[(bi * PB + k) * CLD + j * PB + c];
      Li[(bi * PB + r) * CLD + j * PB + c] = -acc;
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) Mmat[e] = Li[(e / NB) * CLD + (e % NB)];
  if (threadIdx.x == 0) {
    const bool ok = s_min > pivot_tol * s_max;     // NaN compares false -> robust path
    guard->done = ok ? 1u : 0u;
    if (!ok) atomicAdd(&guard->rot[2], 1u);
  }
}

// Polish of a nearly orthonormal block: Mmat = S^-1/2 by the series  (I + D)^-1/2 = I - D/2 + 3 D^2/8 - 5 D^3/16,
// D = S - I, evaluated in shared memory (Horner, two 128^3 products).  The truncation error is (35/128) |D|^4, so the
// guard accepts max|D| < 0.03 (error <= 2e-7); a plain Newton-Schulz step (first two terms) would leave 3/8 |D|^2,
// i.e. 1e-3 at the same threshold -- measured as a 2e-3 loss of orthogonality at p = 603.  Otherwise guard->done = 0
// and the caller's eigen-based SVQB runs.  guard->rot[3] counts those fallbacks.
constexpr int NS_SMEM = 3 * NB * (NB + 1) * 4;
__device__ __forceinline__ void ns_matmul(const float* X, const float* Y, float* out, float cI, float cXY) {
  // out = cI * I + cXY * X Y   (all [NB][NB + 1] in shared memory; 1024 threads, 4 x 4 per thread)
  constexpr int LDS_ = NB + 1;
  const int tr = (threadIdx.x >> 5) * 4, tcn = (threadIdx.x & 31) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k = 0; k < NB; ++k) {
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = X[(tr + i) * LDS_ + k];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = Y[k * LDS_ + tcn + j];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[(tr + i) * LDS_ + tcn + j] = cXY * acc[i][j] + ((tr + i) == (tcn + j) ? cI : 0.f);
}
// `skip`: no-op when *skip != 0 (a later clean-up round of a block that is already finished).  On exit guard->sweeps
// (the "block finished" flag the later rounds test) is 1 when the series was valid, 0 when the SVQB fallback runs.
__global__ void __launch_bounds__(1024) ns_matrix_kernel(const float* __restrict__ S, float* Mmat, bj::Ctl* guard,
                                                         const unsigned int* skip = nullptr) {
  extern __shared__ float nsm[];
  if (skip && *skip) return;
  constexpr int LDS_ = NB + 1;
  float* D = nsm;
  float* E1 = nsm + NB * LDS_;
  float* E2 = nsm + 2 * NB * LDS_;
  __shared__ float red[32];
  float dev = 0.f;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int r = e / NB, c = e % NB;
    const float d = 0.5f * (S[e] + S[c * NB + r]) - (r == c ? 1.f : 0.f);
    dev = fmaxf(dev, fabsf(d));
    D[r * LDS_ + c] = d;
    E1[r * LDS_ + c] = (r == c ? 0.375f : 0.f) - 0.3125f * d;          // 3/8 I - 5/16 D
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dev = fmaxf(dev, __shfl_xor_sync(0xffffffffu, dev, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dev;
  __syncthreads();
  ns_matmul(D, E1, E2, -0.5f, 1.f);                                     // E2 = -1/2 I + D E1
  __syncthreads();
  ns_matmul(D, E2, E1, 1.f, 1.f);                                       // E1 = I + D E2 = S^-1/2 (to 4th order)
  __syncthreads();
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) Mmat[e] = E1[(e / NB) * LDS_ + (e % NB)];
  if (threadIdx.x == 0) {
    float m = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
    const bool ok = m < 0.03f;                  // NaN compares false -> the robust path runs
    guard->done = ok ? 1u : 0u;
    guard->sweeps = ok ? 1u : 0u;
    if (!ok) atomicAdd(&guard->rot[3], 1u);
  }
}

// Qc[c][j0 + r] = Qt_block[r][c]  for one 128-row block of Qt (both matrices P x P row-major)
__global__ void transpose_block_kernel(const float* __restrict__ blk, int P, int j0, float* __restrict__ Qc) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) tile[i][threadIdx.x] = blk[(size_t)(r0 + i) * P + c0 + threadIdx.x];
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) Qc[(size_t)(c0 + i) * P + j0 + r0 + threadIdx.x] = tile[threadIdx.x][i];
}

__global__ void latch_kernel(unsigned int* dst, const unsigned int* src) { if (threadIdx.x == 0) *dst = *src; }
__global__ void copy_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n, const unsigned int* skip) {
  if (skip && *skip) return;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) dst[e] = src[e];
}

// T blocks out of the first H of iteration j: Tdiag[j] = sym(H[:, j]), Tsub[j] = H[:, j-1]
__global__ void extract_T_kernel(const float* __restrict__ H, int P, int j, float* Tdiag, float* Tsub) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < NB * NB; e += gridDim.x * blockDim.x) {
    const int r = e / NB, c = e % NB;
    Tdiag[(int64_t)j * NB * NB + e] = 0.5f * (H[(int64_t)r * P + j * NB + c] + H[(int64_t)c * P + j * NB + r]);
    if (j > 0) Tsub[(int64_t)j * NB * NB + e] = H[(int64_t)r * P + (j - 1) * NB + c];
  }
}

// developer knobs (environment, read per call): second orthonormalisation pass, small-eig tolerance / sweeps
inline int knob_i(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
inline float knob_f(const char* name, float dflt) { const char* e = getenv(name); return e ? (float)atof(e) : dflt; }

struct Ctx {
  Plan pl; char* base; cudaStream_t st; int launches; cudaError_t err;
  const bj::Hooks* hooks;
  int polish_ns = 1; float eig_tol = 1e-6f; int eig_sweeps = 8; int reorth3 = 1; int reorth2 = 0; float chol_tol = 1e-4f; int pass1_local = 1; int extra_rounds = 2;
  float* f(size_t off) const { return (float*)(base + off); }
  int64_t* i64(size_t off) const { return (int64_t*)(base + off); }
};

// C[M x N] (ld ldc) = A(M x K) * B(N x K)^T with automatic split-K for skinny outputs
inline bool gemm(Ctx& cx, const GatherOperand& A, const GatherOperand& B, int M, int N, int K, float* C, int64_t ldc,
                 const unsigned int* skip = nullptr, bool subtract = false, const uint8_t* b_planes = nullptr) {
  const int bn = N > 128 ? 256 : 128;
  int max_splits = (int)std::min<size_t>(64, cx.pl.part_elems / ((size_t)M * N));
  if (max_splits < 1) max_splits = 1;
  const int tiles = ((M + 127) / 128) * ((N + bn - 1) / bn);
  const int splits = tc_choose_splits(tiles, (K + tc::BK - 1) / tc::BK, max_splits, (int64_t)M * N);
  TcGemmParams g = tc_gemm_defaults();
  g.A = A; g.B = B; g.M = M; g.N = N; g.K = K; g.skip_flag = skip;
  if (b_planes && bn == 256) {
    g.b_planes = b_planes; g.b_planes_nkb = (cx.pl.P + tc::BK - 1) / tc::BK; g.b_planes_raw = cx.pl.planes_raw ? 1 : 0;
    static const bool hint_on = [] { const char* e = getenv("MVB_LB_STREAM_HINTS"); return !e || atoi(e) != 0; }();
    g.b_stream_once = (hint_on && M <= 128) ? 1 : 0;      // one M tile: B streams once, the 128-row A block is what every CTA re-reads
  }
  const bool via_parts = splits > 1 || subtract;      // C -= A B^T always goes through the partial-sum buffer
  if (via_parts) { g.C = cx.f(cx.pl.off_part); g.ldc = N; g.c_split_stride = (int64_t)M * N; }
  else { g.C = C; g.ldc = ldc; }
  const bool timed = cx.hooks && !skip;       // conditional (clean-up round) launches are normally no-ops: not profiled
  const int slot = timed ? cx.hooks->begin(2, 2.0 * M * (double)N * K, cx.st) : -1;
  cx.err = tc_gemm_launch(g, splits, bn, cx.st);
  if (timed) cx.hooks->end(slot, cx.st);
  ++cx.launches;
  if (cx.err != cudaSuccess) return false;
  if (via_parts) {
    const int rgrid = (int)std::min<int64_t>(148 * 4, ((int64_t)M * N + 255) / 256);      // small outputs: a few CTAs, leave the SMs to concurrent work
    reduce_parts_kernel<<<rgrid, 256, 0, cx.st>>>(cx.f(cx.pl.off_part), splits, (int64_t)M * N, M, N, C, ldc, skip, subtract ? 1 : 0);
    ++cx.launches;
  }
  return true;
}

// Orthonormalise the 128 rows of `src` (ld P) into `dst` (ld P):  dst = diag(lam^-1/2) J src,  J,lam = eig(src src^T)
inline bool svqb(Ctx& cx, const float* src, float* dst) {
  const Plan& pl = cx.pl;
  const int P = pl.P;
  int64_t *rowP = cx.i64(pl.off_rowP), *row128 = cx.i64(pl.off_row128), *iota = cx.i64(pl.off_iota);
  float *S = cx.f(pl.off_S), *J = cx.f(pl.off_J), *Mm = cx.f(pl.off_M);
  bj::Ctl* ctl = (bj::Ctl*)(cx.base + pl.off_ctl);
  GatherOperand a{src, rowP, iota, NB, 2};
  if (!gemm(cx, a, a, NB, NB, P, S, NB)) return false;
  if (cx.chol_tol > 0.f) {
    // Cholesky QR when the block is well conditioned; otherwise the two kernels below do the eigen-based SVQB
    chol_orth_kernel<<<1, 1024, CHOL_ORTH_SMEM, cx.st>>>(S, Mm, cx.chol_tol, ctl); ++cx.launches;
    bj::small_eig_kernel<<<1, 1024, bj::SMALL_EIG_SMEM, cx.st>>>(S, J, cx.eig_tol, 1.f, cx.eig_sweeps, ctl, 0, cx.f(pl.off_lam) + 8); ++cx.launches;
    svqb_scale_kernel<<<1, 1024, 0, cx.st>>>(J, cx.f(pl.off_lam) + 8, Mm, ctl); ++cx.launches;
  } else {
    bj::small_eig_kernel<<<1, 1024, bj::SMALL_EIG_SMEM, cx.st>>>(S, J, cx.eig_tol, 1.f, cx.eig_sweeps, ctl, 0, cx.f(pl.off_lam) + 8); ++cx.launches;
    svqb_scale_kernel<<<1, 1024, 0, cx.st>>>(J, cx.f(pl.off_lam) + 8, Mm, nullptr); ++cx.launches;
  }
  GatherOperand m{Mm, row128, iota, NB, 2};
  GatherOperand w{src, iota, rowP, P, 0};          // rows = columns c of src, contraction over its 128 rows
  return gemm(cx, m, w, NB, P, NB, dst, P);
}

// second orthonormalisation pass for rows that are already nearly orthonormal: dst = (src src^T)^-1/2 src (series)
inline bool ns_polish(Ctx& cx, const float* src, float* dst, const unsigned int* skip = nullptr) {
  if (!cx.polish_ns) return svqb(cx, src, dst);
  const Plan& pl = cx.pl;
  const int P = pl.P;
  int64_t *rowP = cx.i64(pl.off_rowP), *row128 = cx.i64(pl.off_row128), *iota = cx.i64(pl.off_iota);
  float *S = cx.f(pl.off_S), *J = cx.f(pl.off_J), *Mm = cx.f(pl.off_M);
  bj::Ctl* ctl = (bj::Ctl*)(cx.base + pl.off_ctl);
  GatherOperand a{src, rowP, iota, NB, 2};
  if (!gemm(cx, a, a, NB, NB, P, S, NB, skip)) return false;
  ns_matrix_kernel<<<1, 1024, NS_SMEM, cx.st>>>(S, Mm, ctl, skip); ++cx.launches;
  // robust fallback (rank-deficient residual blocks): both kernels are no-ops while ctl->done is set
  bj::small_eig_kernel<<<1, 1024, bj::SMALL_EIG_SMEM, cx.st>>>(S, J, cx.eig_tol, 1.f, cx.eig_sweeps, ctl, 0, cx.f(pl.off_lam) + 8, skip); ++cx.launches;
  svqb_scale_kernel<<<1, 1024, 0, cx.st>>>(J, cx.f(pl.off_lam) + 8, Mm, ctl, skip); ++cx.launches;
  GatherOperand m{Mm, row128, iota, NB, 2};
  GatherOperand w{src, iota, rowP, P, 0};
  return gemm(cx, m, w, NB, P, NB, dst, P, skip);
}

// W (128 x P) -= (W Q^T) Q  against the first nq rows of Qt
inline bool reorth(Ctx& cx, float* W, const float* Qt, int nq, const unsigned int* skip = nullptr) {
  const Plan& pl = cx.pl;
  const int P = pl.P;
  int64_t *rowP = cx.i64(pl.off_rowP), *iota = cx.i64(pl.off_iota);
  float *H = cx.f(pl.off_H), *W2 = cx.f(pl.off_Wt2);
  const uint8_t* plQt = pl.planes_bytes ? (const uint8_t*)(cx.base + pl.off_plQt) : nullptr;
  const uint8_t* plQc = pl.planes_bytes ? (const uint8_t*)(cx.base + pl.off_plQc) : nullptr;
  GatherOperand w{W, rowP, iota, NB, 2}, q{Qt, rowP, iota, nq, 2};
  if (!gemm(cx, w, q, NB, nq, P, H, P, skip, false, plQt)) return false;
  GatherOperand h{H, rowP, iota, NB, 2}, qc{cx.f(pl.off_Qc), rowP, iota, P, 2};      // rows c of Q^T, k over Q's first nq rows
  return gemm(cx, h, qc, NB, P, nq, W, P, skip, true, plQc);                          // W -= H Q
}

// Enqueue the reduction.  G: p x p.  Outputs: Qt (P x P), Tdiag / Tsub ([nblk][128][128]; Tsub[j] = T_{j,j-1}).
// Returns #launches or -1 (err set).
inline int reduce(const float* G, int p, float* Qt, float* Tdiag, float* Tsub, void* ws, cudaStream_t st, cudaError_t* err,
                  const bj::Hooks* hooks = nullptr) {
  Ctx cx{make_plan(p), (char*)ws, st, 0, cudaSuccess, hooks};
  cx.polish_ns = knob_i("MVB_LB_POLISH_NS", 1); cx.eig_tol = knob_f("MVB_LB_EIG_TOL", 1e-6f);
  cx.eig_sweeps = knob_i("MVB_LB_EIG_SWEEPS", 8); cx.reorth3 = knob_i("MVB_LB_REORTH3", 1);
  cx.pass1_local = knob_i("MVB_LB_PASS1_LOCAL", 1);
  cx.extra_rounds = cx.polish_ns ? std::min(4, std::max(0, knob_i("MVB_LB_EXTRA_ROUNDS", 2))) : 0;
  cx.reorth2 = knob_i("MVB_LB_REORTH2", 0); cx.chol_tol = knob_f("MVB_LB_CHOL_TOL", 1e-4f);
  const Plan& pl = cx.pl;
  const int P = pl.P, nb = pl.nblk;
  float *Gp = cx.f(pl.off_Gp), *Wt = cx.f(pl.off_Wt), *H = cx.f(pl.off_H);
  int64_t *rowP = cx.i64(pl.off_rowP), *row128 = cx.i64(pl.off_row128), *iota = cx.i64(pl.off_iota);
  bj::Ctl* ctl = (bj::Ctl*)(cx.base + pl.off_ctl);
#define LB_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { *err = e_; return -1; } } while (0)
#define LB_OK(x) do { if (!(x)) { *err = cx.err; return -1; } } while (0)
  static bool attr_set = false;
  if (!attr_set) {
    LB_CK(cudaFuncSetAttribute(bj::small_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bj::SMALL_EIG_SMEM));
    LB_CK(cudaFuncSetAttribute(chol_orth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHOL_ORTH_SMEM));
    LB_CK(cudaFuncSetAttribute(ns_matrix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NS_SMEM));
    attr_set = true;
  }
  LB_CK(cudaMemsetAsync(ctl, 0, sizeof(bj::Ctl), st));
  tables_kernel<<<64, 256, 0, st>>>(P, rowP, row128, iota); ++cx.launches;
  diag_mean_kernel<<<1, 256, 0, st>>>(G, p, cx.f(pl.off_lam)); ++cx.launches;
  pad_gram_kernel<<<(int)std::min<int64_t>(148 * 8, ((int64_t)P * P + 255) / 256), 256, 0, st>>>(G, p, P, Gp, cx.f(pl.off_lam)); ++cx.launches;
  uint8_t *plG = nullptr, *plQt = nullptr, *plQc = nullptr;
  if (pl.planes_bytes) {
    plG = (uint8_t*)(cx.base + pl.off_plG); plQt = (uint8_t*)(cx.base + pl.off_plQt); plQc = (uint8_t*)(cx.base + pl.off_plQc);
    GatherOperand g_all{Gp, rowP, iota, P, 2};
    LB_CK(tc_presplit_launch(g_all, P, plG, st, 0, -1, 0, -1, pl.planes_raw)); ++cx.launches;
    LB_CK(cudaMemsetAsync(plQt, 0, pl.planes_bytes, st));       // blocks arrive one per step; rows / stages not yet there read as zero
    LB_CK(cudaMemsetAsync(plQc, 0, pl.planes_bytes, st));
  }
  // Q_0: orthonormalised pseudo-random block
  random_block_kernel<<<(int)std::min<int64_t>(148 * 2, ((int64_t)NB * P + 255) / 256), 256, 0, st>>>(Wt, (int64_t)NB * P, 0x9e3779b9u); ++cx.launches;
  LB_OK(svqb(cx, Wt, Qt));
  LB_OK(ns_polish(cx, Qt, Wt));
  LB_CK(cudaMemcpyAsync(Qt, Wt, (size_t)NB * P * 4, cudaMemcpyDeviceToDevice, st));
  for (int j = 0; j < nb; ++j) {
    const float* Qj = Qt + (size_t)j * NB * P;
    transpose_block_kernel<<<dim3(P / 32, NB / 32), dim3(32, 8), 0, st>>>(Qj, P, j * NB, cx.f(pl.off_Qc)); ++cx.launches;   // block j is final
    if (pl.planes_bytes) {        // ... and joins the pre-split copies of Qt (128 new rows) and Q^T (128 new contraction columns)
      GatherOperand qt_all{Qt, rowP, iota, P, 2}, qc_all{cx.f(pl.off_Qc), rowP, iota, P, 2};
      LB_CK(tc_presplit_launch(qt_all, P, plQt, st, j * NB, (j + 1) * NB, 0, -1, pl.planes_raw)); ++cx.launches;
      LB_CK(tc_presplit_launch(qc_all, P, plQc, st, 0, P, j * (NB / tc::BK), (j + 1) * (NB / tc::BK), pl.planes_raw)); ++cx.launches;
    }
    // W = Q_j Gp ;  H = W Q_{0..j}^T  -> T_jj, T_{j,j-1}
    GatherOperand qj{Qj, rowP, iota, NB, 2}, g{Gp, rowP, iota, P, 2};
    LB_OK(gemm(cx, qj, g, NB, P, P, Wt, P, nullptr, false, plG));
    const int nq = (j + 1) * NB;
    // first pass: against all previous blocks, or (pass1_local) only against the two blocks the three-term
    // recurrence couples -- the later full pass on the normalised block removes the drift along the older ones
    const int b0 = cx.pass1_local ? std::max(0, j - 1) : 0;       // first block of the first pass
    const int nq1 = (j + 1 - b0) * NB;
    {
      GatherOperand w{Wt, rowP, iota, NB, 2}, q{Qt + (size_t)b0 * NB * P, rowP, iota, nq1, 2};
      LB_OK(gemm(cx, w, q, NB, nq1, P, H + (size_t)b0 * NB, P));
      extract_T_kernel<<<16, 256, 0, st>>>(H, P, j, Tdiag, Tsub); ++cx.launches;
    }
    if (j == nb - 1) break;
    {
      GatherOperand h{H + (size_t)b0 * NB, rowP, iota, NB, 2}, qc{cx.f(pl.off_Qc), rowP, iota + (size_t)b0 * NB, P, 2};
      LB_OK(gemm(cx, h, qc, NB, P, nq1, Wt, P, nullptr, true));                        // W -= T_j,j-1 Q_j-1 + T_jj Q_j
    }
    if (cx.reorth2) LB_OK(reorth(cx, Wt, Qt, nq));
    float* Qn = Qt + (size_t)(j + 1) * NB * P;
    LB_OK(svqb(cx, Wt, Qn));
    if (cx.reorth3) LB_OK(reorth(cx, Qn, Qt, nq));
    LB_OK(ns_polish(cx, Qn, Wt));
    LB_CK(cudaMemcpyAsync(Qn, Wt, (size_t)NB * P * 4, cudaMemcpyDeviceToDevice, st));
    // Clean-up rounds for hard blocks only.  When the polish had to fall back to the eigen-based SVQB (ctl->sweeps == 0:
    // the block was ill conditioned, its weak directions were amplified together with their contamination along the
    // previous blocks), re-orthogonalise and polish again; on ordinary steps every kernel below exits at once.
    for (int r = 0; r < cx.extra_rounds; ++r) {
      unsigned int* rs = &ctl->rot[8 + r];       // this round's skip flag: the "finished" state latched at its start
      latch_kernel<<<1, 32, 0, st>>>(rs, &ctl->sweeps); ++cx.launches;
      LB_OK(reorth(cx, Qn, Qt, nq, rs));
      LB_OK(ns_polish(cx, Qn, Wt, rs));          // its ns_matrix_kernel rewrites ctl->sweeps for the next round
      copy_kernel<<<(int)std::min<int64_t>(148 * 2, ((int64_t)NB * P / 4 + 255) / 256), 256, 0, st>>>(Qn, Wt, (int64_t)NB * P, rs); ++cx.launches;
    }
  }
  LB_CK(cudaGetLastError());
#undef LB_CK
#undef LB_OK
  return cx.launches;
}

}  // namespace lb
}  // namespace mvb
