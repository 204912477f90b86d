// dense_batched.cuh -- B independent dense leave-one-out ridge problems of one shape in a handful of launches.
// Two routes behind mvb_dense_ridge_batched: the tridiagonal route (second half of this file, default) and the Cholesky route
// described first (MVB_DENSE_ROUTE=chol).
//
// Replaces the per-slice estimator loop of _Sliding_torch.fit (mvpy/estimators/sliding.py:575-622) around
// _RidgeCV_torch.fit (mvpy/estimators/ridgecv.py:96-303): per-timepoint decoding (BASELINE configs[4]: 500 slices of a
// 1600 x 306 design per fold) is ~180 tiny dependent launches per slice on the general route.  Here slice is a batch index:
//
//   means_kernel / pack_kernel   column means of every slice; Xc[b][i][c] = X_b[i][c] - mean (centred before multiplying,
//                                ridgecv.py:59-94), slice-major, zero-padded to Pp columns
//   tc_gemm (batch = B)          G_b = Xc_b^T Xc_b on the split-TF32 tensor-core kernel
//   xty_kernel                   Xc_b^T Yc_b
//   chol_inv_kernel              one CTA per (slice, alpha): G_b + alpha I = L L^T and L^-1, both in shared memory (packed
//                                lower triangle, blocked by 16, 4 x 4 register tiles); pivots below pivot_tol * max diag flag
//                                the slice in status[] (the caller refits those on the general route)
//   tc_gemm (batch = B)          leverages h[b][a][i] = |L_a^-1 x_i|^2: Xc_b against the stack of all alphas' triangular
//                                inverses, row squares accumulated in the epilogue, nothing else stored, and the zero half of
//                                every triangular block skipped (TcGemmParams::b_tri_period)
//   beta_kernel                  beta = L^-T (L^-1 X^T y) per (slice, alpha)
//   tc_gemm (batch = B)          fitted values Xc_b beta for all (alpha, target)
//   loss_kernel / select_kernel  LOO residuals r / (1 - h - 1/n), loss (A x F), first-minimum alpha (per target or shared,
//                                ridgecv.py:241-262), coefficients and intercepts
//
// The identity used: with M_a = G + alpha I, leverage_i = x_i^T M_a^-1 x_i = |L_a^-1 x_i|^2 (SURVEY.md 8 a-note with the
// eigendecomposition replaced by one Cholesky factor per alpha, which for p <= 320 is cheaper than any batched eigensolver).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mvb {
namespace db {

constexpr int MAX_COLS = 320;      // packed lower triangle of 320 x 320 floats + one 16-column panel fit in 227 KB
constexpr int NB = 16;             // panel width of the blocked factorisation / inversion
constexpr int CI_THREADS = 512;
constexpr int TRI_TILE = 128;      // N tile of the leverage product: one alpha's block is `tri_rows / 128` tiles

struct View { const float* base; int64_t s_sample, s_col, s_batch; };

__host__ __device__ __forceinline__ int tri(int i) { return i * (i + 1) / 2; }

// out[b * ldo + c] = mean over the selected samples of X_b[., c]   (fp64 accumulation; 0 when !centre)
__global__ void means_kernel(View v, const int32_t* __restrict__ idx, int n, int ncol, int B, int centre, float* __restrict__ out, int ldo) {
  const bool batch_fast = v.s_batch == 1;
  const int64_t total = (int64_t)ncol * B;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int b = batch_fast ? (int)(e % B) : (int)(e / ncol);
    const int c = batch_fast ? (int)(e / B) : (int)(e % ncol);
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    if (centre) {
      const float* src = v.base + (int64_t)b * v.s_batch + (int64_t)c * v.s_col;
      int i = 0;
      for (; i + 3 < n; i += 4) {
        const int64_t r0 = idx ? idx[i] : i, r1 = idx ? idx[i + 1] : i + 1, r2 = idx ? idx[i + 2] : i + 2, r3 = idx ? idx[i + 3] : i + 3;
        s0 += (double)__ldg(src + r0 * v.s_sample); s1 += (double)__ldg(src + r1 * v.s_sample);
        s2 += (double)__ldg(src + r2 * v.s_sample); s3 += (double)__ldg(src + r3 * v.s_sample);
      }
      for (; i < n; ++i) s0 += (double)__ldg(src + (int64_t)(idx ? idx[i] : i) * v.s_sample);
    }
    out[(int64_t)b * ldo + c] = (float)(((s0 + s1) + (s2 + s3)) / (double)n);
  }
}

// the same for few columns (the targets: 3 x 500 at configs[4]): one warp per (slice, column), lanes over the samples -- the
// thread-per-column kernel above would walk 1600 samples with 1500 threads (0.58 ms of latency)
__global__ void means_warp_kernel(View v, const int32_t* __restrict__ idx, int n, int ncol, int B, int centre, float* __restrict__ out, int ldo) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (int64_t)ncol * B) return;
  const int b = (int)(w / ncol), c = (int)(w % ncol);
  double s = 0;
  if (centre) {
    const float* src = v.base + (int64_t)b * v.s_batch + (int64_t)c * v.s_col;
    for (int i = lane; i < n; i += 32) s += (double)__ldg(src + (int64_t)(idx ? idx[i] : i) * v.s_sample);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[(int64_t)b * ldo + c] = (float)(s / (double)n);
}

// out[(b * n + i) * ld + c] = X_b[idx[i], c] - mean[b * ldm + c]  (c < ncol), 0 for the padding columns.
// grid (ceil(ld / 32), ceil(B / 32), n), block (32, 8): a 32 x 32 (column, slice) tile of one sample goes through shared
// memory so that both the strided source (slices contiguous for a Sliding view) and the destination are read / written in rows.
__global__ void pack_kernel(View v, const int32_t* __restrict__ idx, int n, int ncol, int B, const float* __restrict__ mean, int ldm,
                            float* __restrict__ out, int ld) {
  __shared__ float tile[32][33];
  const int i = blockIdx.z, c0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  const int64_t row = idx ? idx[i] : i;
  const float* src = v.base + row * v.s_sample;
  if (v.s_batch == 1) {
    for (int cy = threadIdx.y; cy < 32; cy += blockDim.y) {
      const int c = c0 + cy, b = b0 + threadIdx.x;
      tile[cy][threadIdx.x] = (c < ncol && b < B) ? __ldg(src + (int64_t)c * v.s_col + b) : 0.f;
    }
  } else {
    for (int by = threadIdx.y; by < 32; by += blockDim.y) {
      const int b = b0 + by, c = c0 + threadIdx.x;
      tile[threadIdx.x][by] = (c < ncol && b < B) ? __ldg(src + (int64_t)b * v.s_batch + (int64_t)c * v.s_col) : 0.f;
    }
  }
  __syncthreads();
  for (int by = threadIdx.y; by < 32; by += blockDim.y) {
    const int b = b0 + by, c = c0 + threadIdx.x;
    if (b < B && c < ld)
      out[((int64_t)b * n + i) * ld + c] = (c < ncol) ? tile[threadIdx.x][by] - mean[(int64_t)b * ldm + c] : 0.f;
  }
}

// gather tables of the batched products (element offsets)
__global__ void tables_kernel(int B, int n, int Pp, int AF, int A, int tri_rows, int64_t* roffX, int64_t* roffXt, int64_t* koffN,
                              int64_t* iota, int64_t* roffLi, int64_t* roffBeta) {
  const int64_t nx = (int64_t)B * n, nxt = (int64_t)B * Pp, nli = (int64_t)B * A * tri_rows, nbt = (int64_t)B * AF;
  const int64_t niota = n > Pp ? n : Pp;
  const int64_t total = nx + nxt + n + niota + nli + nbt;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = e;
    if (t < nx) { roffX[t] = t * Pp; continue; }                                   // row i of slice b
    t -= nx;
    if (t < nxt) { roffXt[t] = (t / Pp) * (int64_t)n * Pp + (t % Pp); continue; }  // column c of slice b (rows contiguous)
    t -= nxt;
    if (t < n) { koffN[t] = t * Pp; continue; }
    t -= n;
    if (t < niota) { iota[t] = t; continue; }
    t -= niota;
    if (t < nli) { roffLi[t] = t * Pp; continue; }                                 // row j of block (b, a): [B][A][tri_rows][Pp]
    t -= nli;
    roffBeta[t] = t * Pp;                                                          // row (a, f) of slice b: [B][A][F][Pp]
  }
}

// XtY[(b * Pp + c) * F + f] = sum_i Xc[b][i][c] Yc[b][i][f];  grid (ceil(Pp / 128), B, ceil(F / 8)), block 128
__global__ void xty_kernel(const float* __restrict__ Xc, const float* __restrict__ Yc, int n, int Pp, int F, float* __restrict__ XtY) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y, f0 = blockIdx.z * 8;
  if (c >= Pp) return;
  const int nf = min(8, F - f0);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const float* x = Xc + (int64_t)b * n * Pp + c;
  const float* y = Yc + (int64_t)b * n * F + f0;
  for (int i = 0; i < n; ++i) {
    const float xv = x[(int64_t)i * Pp];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < nf) acc[j] += xv * __ldg(y + (int64_t)i * F + j);
  }
  for (int j = 0; j < nf; ++j) XtY[((int64_t)b * Pp + c) * F + f0 + j] = acc[j];
}

// One CTA per (slice b, alpha a): M = G_b + alpha I (padding rows: identity) -> L (Cholesky) -> L^-1, written as a full
// [tri_rows][Pp] block (zeros above the diagonal and in the rows beyond Pp) for the leverage product.
// Shared memory: the packed lower triangle (row i at tri(i)) + a [16][Pp] k-major copy of the current panel + the 16 x 17
// inverse of the current diagonal block.  What the phases cost was measured with the counters below (tools/chol_inv_phases.py);
// the layout choices follow from it:
//   * 16 x 16 diagonal blocks are factored / inverted in the registers of one warp (lane = row, values exchanged by shuffles):
//     the same steps on shared memory were 27 % of the kernel;
//   * the trailing update reads the panel from the k-major copy -- for one k the 8 rows of a tile are two 16-byte loads and the
//     lanes of a warp (consecutive column pairs of one row group) read consecutive 8-byte words, no bank conflicts -- instead
//     of from the packed rows, whose start offsets tri(i) scatter the lanes over the banks (3-way conflicts, 26 %);
//   * the inverse's block-column update splits every tile's k range over 4 adjacent lanes (all 512 threads busy instead of
//     m <= 304 with trip counts up to m).
inline size_t chol_inv_smem(int Pp) { return ((size_t)((tri(Pp) + 3) & ~3) + (size_t)Pp * NB + NB * (NB + 1)) * 4; }

// lane r (mod 16) loads row r of the 16 x 16 lower-triangular diagonal block at J into registers
__device__ __forceinline__ void diag_load(const float* __restrict__ Lp, int J, int r, float (&row)[NB]) {
#pragma unroll
  for (int c = 0; c < NB; ++c) row[c] = (c <= r) ? Lp[tri(J + r) + J + c] : 0.f;
}

// in-register Cholesky of the diagonal block: on exit lane r holds row r of L.  Returns true when a pivot of a real
// (non-padding) row was not above floor_.
__device__ __forceinline__ bool diag_factor(float (&row)[NB], int r, int n_real, float floor_) {
  bool bad = false;
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    float piv = __shfl_sync(0xffffffffu, row[k], k);
    if (k < n_real && !(piv > floor_)) { bad = true; piv = floor_; }      // (padding rows are an identity block: pivot 1)
    // 1 / sqrt(piv) from the special-function unit + one Newton step (full fp32 accuracy): the precise sqrtf and division
    // are ~80 dependent instructions per column on the one warp that works here
    float is = rsqrtf(piv);
    is = is * (1.5f - 0.5f * piv * is * is);
    if (r == k) row[k] = piv * is;
    else if (r > k) row[k] = row[k] * is;
    const float lrk = row[k];
#pragma unroll
    for (int c = k + 1; c < NB; ++c) {
      const float lck = __shfl_sync(0xffffffffu, row[k], c);              // L[c][k], scaled just above by lane c
      if (r >= c) row[c] -= lrk * lck;
    }
  }
  return bad;
}

// inverse of the lower-triangular block held row-wise in registers: lane c (mod 16) solves column c by forward substitution
// (L[i][k] comes from lane i by shuffle) and writes it to Dinv[i][c]
__device__ __forceinline__ void diag_invert(const float (&row)[NB], int lane, float* __restrict__ Dinv) {
  const int c = lane & (NB - 1);
  float x[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    float v = (i == c) ? 1.f : 0.f;
#pragma unroll
    for (int k = 0; k < NB; ++k)
      if (k < i) {
        const float lik = __shfl_sync(0xffffffffu, row[k], i);
        if (k >= c) v -= lik * x[k];
      }
    const float lii = __shfl_sync(0xffffffffu, row[i], i);
    x[i] = (i < c) ? 0.f : v * __frcp_rn(lii);
  }
  if (lane < NB) {
#pragma unroll
    for (int i = 0; i < NB; ++i) Dinv[i * (NB + 1) + c] = x[i];
  }
}

// The two warp-level steps as separate (not inlined) functions: inlined into the 128-register kernel the compiler kept the
// 16-value row in local memory (304 LDL / STL, 12 000 cycles per block instead of ~1 500; cuobjdump -sass).
__device__ __noinline__ bool diag_block_factor(float* __restrict__ Lp, int kb, int n_real, float floor_, float* __restrict__ rdiag, int lane) {
  const int r = lane & (NB - 1);
  float row[NB];
  diag_load(Lp, kb, r, row);
  const bool bad = diag_factor(row, r, n_real, floor_);
  if (lane < NB) {
#pragma unroll
    for (int c = 0; c < NB; ++c)
      if (c <= r) Lp[tri(kb + r) + kb + c] = row[c];
#pragma unroll
    for (int c = 0; c < NB; ++c)
      if (c == r) rdiag[c] = __frcp_rn(row[c]);               // reciprocal diagonal for the panel solve
  }
  return bad;
}
// in place: the lower-triangular diagonal block at J is replaced by its inverse (lane c solves column c from the rows held in
// the warp's registers, so every load precedes every store)
__device__ __noinline__ void diag_block_invert_inplace(float* __restrict__ Lp, int J, int lane) {
  float row[NB];
  diag_load(Lp, J, lane & (NB - 1), row);
  const int c = lane & (NB - 1);
  float x[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    float v = (i == c) ? 1.f : 0.f;
#pragma unroll
    for (int k = 0; k < NB; ++k)
      if (k < i) {
        const float lik = __shfl_sync(0xffffffffu, row[k], i);
        if (k >= c) v -= lik * x[k];
      }
    const float lii = __shfl_sync(0xffffffffu, row[i], i);
    x[i] = (i < c) ? 0.f : v * __frcp_rn(lii);
  }
  __syncwarp();
  if (lane < NB) {
#pragma unroll
    for (int i = 0; i < NB; ++i)
      if (i >= c) Lp[tri(J + i) + J + c] = x[i];
  }
}

__global__ void __launch_bounds__(CI_THREADS, 1) chol_inv_kernel(const float* __restrict__ G, int p, int Pp, const float* __restrict__ alphas,
                                                                int A, float pivot_tol, float* __restrict__ Li, int tri_rows,
                                                                int* __restrict__ status, long long* __restrict__ phase_clk = nullptr) {
  extern __shared__ __align__(16) float sm[];
  float* Lp = sm;
  float* panelT = sm + ((tri(Pp) + 3) & ~3);       // [16][Pp]: panelT[k * Pp + i] = panel entry (row base + i, column k)
  float* Dinv = panelT + Pp * NB;
  __shared__ unsigned int s_dmax;
  __shared__ int s_bad;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int prob = blockIdx.x, b = prob / A, a = prob - b * A;
  const float alpha = alphas[a];
  const float* Gb = G + (int64_t)b * Pp * Pp;
  if (tid == 0) { s_dmax = 0u; s_bad = 0; }
  __syncthreads();
  // developer timing (tools/chol_inv_phases.py): cycles of block 0 per phase, [0] load [1] diag factor [2] panel solve
  // [3] trailing update [4] diag inverse [5] panel product [6] inverse update [7] store
  long long t_last = phase_clk ? clock64() : 0;
  auto tick = [&](int ph) {
    if (phase_clk && prob == 0 && tid == 0) { const long long t = clock64(); phase_clk[ph] += t - t_last; t_last = t; }
  };
  float dmax = 0.f;
  for (int i = warp; i < Pp; i += CI_THREADS / 32) {          // one warp per row of the lower triangle
    for (int j = lane; j <= i; j += 32) {
      float v = (i < p) ? Gb[(int64_t)i * Pp + j] : 0.f;
      if (i == j) { v = (i < p) ? v + alpha : 1.f; if (i < p) dmax = fmaxf(dmax, v); }
      Lp[tri(i) + j] = v;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  if (lane == 0 && dmax > 0.f) atomicMax(&s_dmax, __float_as_uint(dmax));
  __syncthreads();
  const float floor_ = fmaxf(__uint_as_float(s_dmax) * pivot_tol, 1e-30f);
  tick(0);

  // ---- blocked right-looking Cholesky ----
  for (int kb = 0; kb < Pp; kb += NB) {
    if (warp == 0) {                               // (a) the 16 x 16 diagonal block, in registers
      const bool bad = diag_block_factor(Lp, kb, p - kb, floor_, Dinv, lane);
      if (bad && lane == 0) s_bad = 1;
    }
    __syncthreads();
    tick(1);
    const int base = kb + NB, m = Pp - base;
    for (int i = base + tid; i < Pp; i += CI_THREADS) {         // (b) rows below: L[i, panel] = A[i, panel] L_kk^-T
      float* row = Lp + tri(i) + kb;
      float x[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        float v = row[c];
#pragma unroll
        for (int k = 0; k < NB; ++k)
          if (k < c) v -= x[k] * Lp[tri(kb + c) + kb + k];
        x[c] = v * Dinv[c];
      }
#pragma unroll
      for (int c = 0; c < NB; ++c) { row[c] = x[c]; panelT[c * Pp + (i - base)] = x[c]; }
    }
    __syncthreads();
    tick(2);
    {                                                          // (c) trailing update, 8 x 2 register tiles of the lower triangle
      const int nrg = m / 8, ntiles = 2 * nrg * (nrg + 1);     // row group ti (8 rows) has 4 (ti + 1) column pairs
      for (int e = tid; e < ntiles; e += CI_THREADS) {
        int ti = (int)((sqrtf(2.f * (float)e + 1.f) - 1.f) * 0.5f);
        while (2 * (ti + 1) * (ti + 2) <= e) ++ti;
        while (2 * ti * (ti + 1) > e) --ti;
        const int tj = e - 2 * ti * (ti + 1);
        const int i0 = 8 * ti, j0 = 2 * tj;                    // relative to base
        float acc[8][2];
#pragma unroll
        for (int r = 0; r < 8; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; }
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          const float4 a0 = *reinterpret_cast<const float4*>(panelT + k * Pp + i0);
          const float4 a1 = *reinterpret_cast<const float4*>(panelT + k * Pp + i0 + 4);
          const float2 bv = *reinterpret_cast<const float2*>(panelT + k * Pp + j0);
          const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
          for (int r = 0; r < 8; ++r) { acc[r][0] += av[r] * bv.x; acc[r][1] += av[r] * bv.y; }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          float* dst = Lp + tri(base + i0 + r) + base + j0;
          if (j0 <= i0 + r) dst[0] -= acc[r][0];
          if (j0 + 1 <= i0 + r) dst[1] -= acc[r][1];
        }
      }
    }
    __syncthreads();
    tick(3);
  }

  // ---- in-place inverse of the triangular factor ----
  // all diagonal blocks first (independent of each other once the factor is complete: 16 warps, one block each, instead of one
  // serial one-warp phase per block column), then the block columns from the right
  for (int blk = warp; blk < Pp / NB; blk += CI_THREADS / 32) diag_block_invert_inplace(Lp, blk * NB, lane);
  __syncthreads();
  tick(4);
  for (int J = Pp - NB; J >= 0; J -= NB) {
    const int base = J + NB, m = Pp - base;
    float* panel = panelT;                                       // here row-major [m][16]: panel[(i - base) * 16 + c]
    for (int i = base + tid; i < Pp; i += CI_THREADS) {          // panel[i] = L[i, J..J+16) Dinv
      const float* row = Lp + tri(i) + J;
      float l[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) l[k] = row[k];
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < NB; ++k)
          if (k >= c) v += l[k] * Lp[tri(J + k) + J + c];        // the block's inverse, already in place (same address for all threads)
        panel[(i - base) * NB + c] = v;
      }
    }
    __syncthreads();
    tick(5);
    {                                                            // Inv[R, J] = -Inv[R, R] panel[R]   (R = rows below the block)
      // tile = 4 rows x 4 of the 16 columns; its k range is split over 4 adjacent lanes (k = q mod 4) and summed by shuffles
      const int nrt = m / 4, total = nrt * 4 * 4;
      for (int wb = 0; wb < total; wb += CI_THREADS) {
        const int w = wb + tid;
        const bool valid = w < total;
        const int t = valid ? (w >> 2) : 0, q = w & 3;
        const int ti = nrt - 1 - (t >> 2), c0 = (t & 3) * 4;     // long rows first
        const int i0 = 4 * ti;                                   // relative to base
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
        if (valid) {
          const float* ra[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) ra[r] = Lp + tri(base + i0 + r) + base;
          for (int k = q; k < i0; k += 4) {                      // full part: k below the tile's first row
            const float4 tv = *reinterpret_cast<const float4*>(panel + k * NB + c0);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const float av = ra[r][k];
              acc[r][0] += av * tv.x; acc[r][1] += av * tv.y; acc[r][2] += av * tv.z; acc[r][3] += av * tv.w;
            }
          }
          {                                                      // triangular corner: lane q takes k = i0 + q
            const int k = i0 + q;
            const float4 tv = *reinterpret_cast<const float4*>(panel + k * NB + c0);
#pragma unroll
            for (int r = 0; r < 4; ++r)
              if (q <= r) {
                const float av = ra[r][k];
                acc[r][0] += av * tv.x; acc[r][1] += av * tv.y; acc[r][2] += av * tv.z; acc[r][3] += av * tv.w;
              }
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            acc[r][c] += __shfl_xor_sync(0xffffffffu, acc[r][c], 1);
            acc[r][c] += __shfl_xor_sync(0xffffffffu, acc[r][c], 2);
          }
        if (valid) {                                             // lane q writes row q of the tile
          float* dst = Lp + tri(base + i0 + q) + J + c0;
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if (r == q) { dst[0] = -acc[r][0]; dst[1] = -acc[r][1]; dst[2] = -acc[r][2]; dst[3] = -acc[r][3]; }
        }
      }
    }
    __syncthreads();
    tick(6);
  }

  float* out = Li + (int64_t)prob * tri_rows * Pp;
  for (int e = tid; e < tri_rows * Pp; e += CI_THREADS) {
    const int i = e / Pp, j = e - i * Pp;
    out[e] = (i < Pp && j <= i) ? Lp[tri(i) + j] : 0.f;
  }
  tick(7);
  if (tid == 0 && s_bad) atomicOr(&status[b], 1);
}

// beta[(prob * F + f) * Pp + k] = sum_{j >= k} Li[j][k] t[j][f],  t[j][f] = sum_{k <= j} Li[j][k] XtY[b][k][f]
// one CTA (256 threads) per (slice, alpha); targets in chunks of 4
__global__ void __launch_bounds__(256) beta_kernel(const float* __restrict__ Li, int tri_rows, int Pp, const float* __restrict__ XtY, int F,
                                                   int A, float* __restrict__ beta) {
  __shared__ float sx[MAX_COLS][4];
  __shared__ float st[MAX_COLS][4];
  const int prob = blockIdx.x, b = prob / A;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* L = Li + (int64_t)prob * tri_rows * Pp;
  for (int f0 = 0; f0 < F; f0 += 4) {
    const int nf = min(4, F - f0);
    for (int e = tid; e < Pp * 4; e += 256) {
      const int k = e >> 2, j = e & 3;
      sx[k][j] = (j < nf) ? XtY[((int64_t)b * Pp + k) * F + f0 + j] : 0.f;
    }
    __syncthreads();
    for (int j = warp; j < Pp; j += 8) {                   // t[j] = Li[j, 0..j] . XtY
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      for (int k = lane; k <= j; k += 32) {
        const float l = L[(int64_t)j * Pp + k];
        s0 += l * sx[k][0]; s1 += l * sx[k][1]; s2 += l * sx[k][2]; s3 += l * sx[k][3];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
      }
      if (lane == 0) { st[j][0] = s0; st[j][1] = s1; st[j][2] = s2; st[j][3] = s3; }
    }
    __syncthreads();
    for (int k = tid; k < Pp; k += 256) {                  // beta[k] = sum_{j >= k} Li[j][k] t[j]
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      for (int j = k; j < Pp; ++j) {
        const float l = L[(int64_t)j * Pp + k];
        s0 += l * st[j][0]; s1 += l * st[j][1]; s2 += l * st[j][2]; s3 += l * st[j][3];
      }
      const float s[4] = {s0, s1, s2, s3};
      for (int j = 0; j < nf; ++j) beta[((int64_t)prob * F + f0 + j) * Pp + k] = s[j];
    }
    __syncthreads();
  }
}

// d[(b * A + a) * n + i] = 1 - c0 - sum over the alpha's N tiles and the 4 column parts of the row squares
__global__ void leverage_d_kernel(const float* __restrict__ rowsq, int B, int A, int n, int tiles_per_alpha, float c0, float* __restrict__ d) {
  const int64_t total = (int64_t)B * A * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e % n);
    const int64_t ba = e / n;
    const int a = (int)(ba % A), b = (int)(ba / A);
    float h = 0.f;
    for (int t = 0; t < tiles_per_alpha; ++t)
#pragma unroll
      for (int part = 0; part < 4; ++part)
        h += rowsq[((((int64_t)a * tiles_per_alpha + t) * 4 + part) * B + b) * n + i];
    d[e] = 1.f - h - c0;
  }
}

// partial[(b * slabs + s) * AF + af] = sum over the slab's rows of ((Yc[b][i][f] - R[b][i][af]) / d[b][a][i])^2
// grid (B, slabs), block 128 (threads over af)
__global__ void loss_kernel(const float* __restrict__ R, const float* __restrict__ Yc, const float* __restrict__ d, int n, int A, int F,
                            int slabs, double* __restrict__ partial) {
  const int b = blockIdx.x, s = blockIdx.y, AF = A * F;
  const int i0 = (int)((int64_t)n * s / slabs), i1 = (int)((int64_t)n * (s + 1) / slabs);
  for (int af = threadIdx.x; af < AF; af += blockDim.x) {
    const int a = af / F, f = af - a * F;
    const float* r = R + (int64_t)b * n * AF + af;
    const float* y = Yc + (int64_t)b * n * F + f;
    const float* dd = d + ((int64_t)b * A + a) * n;
    double acc = 0;
    float run = 0.f;
    for (int i = i0; i < i1; ++i) {
      const float e = (y[(int64_t)i * F] - r[(int64_t)i * AF]) / dd[i];
      run += e * e;
      if (((i - i0) & 31) == 31) { acc += (double)run; run = 0.f; }
    }
    partial[((int64_t)b * slabs + s) * AF + af] = acc + (double)run;
  }
}

// per slice: loss (A x F), first-minimum alpha (same rules as select_alpha_kernel: a non-finite loss never wins over a finite
// one; shared alpha = argmin of the fp32 mean over targets), coefficients and intercepts.  grid B, block 256.
__global__ void select_kernel(const double* __restrict__ partial, int slabs, int n, int A, int F, int p, int Pp, const float* __restrict__ alphas,
                              int per_target, int fit_intercept, const float* __restrict__ beta, const float* __restrict__ mu,
                              const float* __restrict__ ybar, float* __restrict__ loss_out, int32_t* __restrict__ best, float* __restrict__ alpha_out,
                              float* __restrict__ coef, float* __restrict__ intercept) {
  extern __shared__ float sl[];          // [A * F] losses, then [F] best indices (as int)
  int* sbest = reinterpret_cast<int*>(sl + A * F);
  const int b = blockIdx.x, AF = A * F, tid = threadIdx.x;
  for (int c = tid; c < AF; c += blockDim.x) {
    double s = 0;
    for (int z = 0; z < slabs; ++z) s += partial[((int64_t)b * slabs + z) * AF + c];
    const float v = (float)(s / (double)n);
    sl[c] = v;
    if (loss_out) loss_out[(int64_t)b * AF + c] = v;
  }
  __syncthreads();
  if (per_target) {
    for (int f = tid; f < F; f += blockDim.x) {
      int bi = 0; float bv = sl[f];
      for (int a = 1; a < A; ++a) {
        const float v = sl[a * F + f];
        if (v < bv || (!(bv == bv) && v == v)) { bv = v; bi = a; }
      }
      sbest[f] = bi;
      best[(int64_t)b * F + f] = bi;
      alpha_out[(int64_t)b * F + f] = alphas[bi];
    }
  } else if (tid == 0) {
    int bi = 0; float bv = 0.f;
    for (int a = 0; a < A; ++a) {
      double s = 0;
      for (int f = 0; f < F; ++f) s += (double)sl[a * F + f];
      const float v = (float)(s / F);
      if (a == 0 || v < bv || (!(bv == bv) && v == v)) { bv = v; bi = a; }
    }
    for (int f = 0; f < F; ++f) sbest[f] = bi;
    best[b] = bi;
    alpha_out[b] = alphas[bi];
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  for (int f = warp; f < F; f += nw) {
    const float* src = beta + (((int64_t)b * A + sbest[f]) * F + f) * Pp;
    double m = 0;
    for (int c = lane; c < p; c += 32) {
      const float v = src[c];
      coef[((int64_t)b * F + f) * p + c] = v;
      if (mu) m += (double)v * (double)mu[(int64_t)b * Pp + c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
    if (lane == 0 && mu) intercept[(int64_t)b * F + f] = fit_intercept ? (float)((double)ybar[(int64_t)b * F + f] - m) : 0.f;
  }
}

// =====================================================================================================================
// Tridiagonal route (MVB_DENSE_ROUTE=tridiag, default for the shapes it takes): ONE orthogonal reduction per slice instead of one
// Cholesky factor + inverse + leverage product per (slice, alpha).
//
//   G_b = Q T Q^T   (Householder, T tridiagonal; tridiag_kernel, one CTA per slice, packed lower triangle in shared memory)
//   Z_b = Xc_b Q    (the p - 2 reflectors applied to 32-row tiles held in shared memory; rows_apply_leverage_kernel)
//   h_i(alpha) = z_i^T (T + alpha I)^-1 z_i  from the L D L^T factors of the tridiagonal matrices (tri_ldl_solve_kernel):
//                w_k = z_k - l_{k-1} w_{k-1},  h = sum_k w_k^2 / d_k       -- O(p) per (row, alpha), in the same kernel
//   x_alpha = (T + alpha I)^-1 Q^T X^T y  (tridiagonal solves), fitted values Z x_alpha (batched tensor-core product, as before),
//   coefficients of the chosen alphas back in the original basis: coef = Q x   (reflectors_small_kernel, reverse order)
//
// This is the eigen-route of SURVEY 8 a-note with "eigendecomposition" replaced by the tridiagonal form, which is what a
// symmetric eigensolver computes first anyway and all that the alpha sweep needs: 2 p^3 / 3 + n p^2 fp32 FMAs per slice instead
// of 20 x (2 p^3 / 3 + n p^2).
// =====================================================================================================================
constexpr int TD_THREADS = 1024;
constexpr int TD_NC = MAX_COLS / 32;       // 32-column chunks of a row: lane l of a warp owns columns 32 c + l
inline size_t tridiag_smem(int p) { return ((size_t)((tri(p) + 3) & ~3) + 11 * (size_t)((p + 3) & ~3) + 64) * 4; }

// block sum, all threads get the result: warp shuffles, then warp 0 adds the 32 warp sums (three barriers, two shared-memory reads
// per thread; a first version in which every thread added the 32 partial sums itself spent 2 000 cycles per call on those reads)
__device__ __forceinline__ float block_sum_1024(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) red[32] = s;
  }
  __syncthreads();
  const float s = red[32];
  __syncthreads();
  return s;
}

// dT [B][ldt], eT [B][ldt] (e[k] = T[k+1][k]; the last slot holds the largest diagonal entry of G), V [B][p][ldv] (row k = v_k: 0 up to k,
// 1 at k + 1, then the reflector), tau [B][ldt].
// Per step the trailing block is read ONCE for the symmetric product and once for the rank-2 update, row by row: a warp owns rows
// i = warp (mod 32), its lanes the columns j = lane (mod 32) -- contiguous, conflict-free -- and one packed element A[i][j] serves both
// y_i += A_ij v_j (row sum, reduced by shuffles) and y_j += A_ij v_i (kept per lane in registers, the 32 warps' partial vectors are
// combined through an 8-row buffer in four fixed rounds: deterministic).  First version (four lanes per row walking the row and
// the column of the packed triangle): 16 000 cycles per step, 8.5 ms per fold.
__device__ long long g_td_clk[8];       // developer timing: cycles of block 0 per phase, summed over the steps (tools/tri_phases.py)
__global__ void __launch_bounds__(TD_THREADS, 1) tridiag_kernel(const float* __restrict__ G, int p, int Pp, float* __restrict__ dT,
                                                                float* __restrict__ eT, int ldt, float* __restrict__ V, int ldv,
                                                                float* __restrict__ tau_out) {
  extern __shared__ __align__(16) float sm[];
  const int pa = (p + 3) & ~3;
  float* Lp = sm;
  float* sv = sm + ((tri(p) + 3) & ~3);
  float* sw = sv + pa;
  float* sp = sw + pa;
  float* part = sp + pa;           // [8][pa]
  float* red = part + 8 * pa;      // 32 + scalars
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, b = blockIdx.x;
  const float* Gb = G + (int64_t)b * Pp * Pp;
  float* Vb = V + (int64_t)b * p * ldv;
  float gpart = 0.f;
  for (int i = warp; i < p; i += TD_THREADS / 32)
    for (int j = lane; j <= i; j += 32) {
      const float v = Gb[(int64_t)i * Pp + j];
      Lp[tri(i) + j] = v;
      if (i == j) gpart = fmaxf(gpart, v);
    }
  for (int e = tid; e < p * ldv; e += TD_THREADS) Vb[e] = 0.f;
  // largest diagonal entry of G: the scale of the pivot test in tri_ldl_solve_kernel (the same scale the Cholesky route uses), kept in the
  // unused last slot of e
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gpart = fmaxf(gpart, __shfl_xor_sync(0xffffffffu, gpart, o));
  if (lane == 0) red[warp] = gpart;
  __syncthreads();
  if (tid == 0) {
    float gmax = 0.f;
    for (int w = 0; w < TD_THREADS / 32; ++w) gmax = fmaxf(gmax, red[w]);
    red[40] = gmax;
  }
  __syncthreads();
  const float gmax_all = red[40];
  const bool timed = blockIdx.x == 0 && tid == 0;
  long long t_last = timed ? clock64() : 0, acc_clk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  auto tick = [&](int ph) { if (timed) { const long long t = clock64(); acc_clk[ph] += t - t_last; t_last = t; } };
  for (int k = 0; k + 2 < p; ++k) {
    const int o = k + 1, m = p - o;
    // column k below the diagonal
    float ssq = 0.f;
    for (int j = tid; j < m; j += TD_THREADS) {
      const float x = Lp[tri(o + j) + k];
      sv[j] = x;
      if (j >= 1) ssq += x * x;
    }
    const float xnorm2 = block_sum_1024(ssq, red);           // (its barriers also publish sv)
    tick(0);
    const float alpha0 = sv[0];
    float tau = 0.f, beta = alpha0, scale = 0.f;
    if (xnorm2 > 0.f) {
      beta = -copysignf(sqrtf(alpha0 * alpha0 + xnorm2), alpha0);
      tau = (beta - alpha0) / beta;
      scale = 1.f / (alpha0 - beta);
    }
    __syncthreads();                                           // everybody has read sv[0]
    for (int j = tid; j < m; j += TD_THREADS) {
      const float v = (j == 0) ? 1.f : sv[j] * scale;
      sv[j] = v;
      Vb[(int64_t)k * ldv + o + j] = v;
    }
    if (tid == 0) {
      dT[(int64_t)b * ldt + k] = Lp[tri(k) + k];
      eT[(int64_t)b * ldt + k] = beta;
      tau_out[(int64_t)b * ldt + k] = tau;
    }
    __syncthreads();
    tick(1);
    if (tau != 0.f) {                                          // (uniform: every thread computed the same tau)
      float vr[TD_NC], yc[TD_NC];
#pragma unroll
      for (int c = 0; c < TD_NC; ++c) { const int j = 32 * c + lane; vr[c] = (j < m) ? sv[j] : 0.f; yc[c] = 0.f; }
      // ---- y = A22 v: one pass over the packed rows ----
      for (int i = warp; i < m; i += TD_THREADS / 32) {
        const float* row = Lp + tri(o + i) + o;
        const float vi = sv[i];
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < TD_NC; ++c) {
          if (32 * c > i) break;                               // (uniform per warp)
          const int j = 32 * c + lane;
          if (j <= i) {
            const float a = row[j];
            dot += a * vr[c];                                  // row part (includes the diagonal)
            if (j < i) yc[c] += a * vi;                        // the mirrored element: column part of y_j
          }
        }
#pragma unroll
        for (int s2 = 16; s2 > 0; s2 >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, s2);
        if (lane == 0) sp[i] = dot;
      }
      // the 32 warps' column parts, eight at a time into an 8-row buffer (fixed order)
      for (int round = 0; round < 4; ++round) {
        __syncthreads();
        if (round == 0) tick(2);
        if ((warp >> 3) == round) {
          float* dst = part + (warp & 7) * pa;
#pragma unroll
          for (int c = 0; c < TD_NC; ++c) {
            const int j = 32 * c + lane;
            if (j < m) dst[j] = (round == 0) ? yc[c] : dst[j] + yc[c];
          }
        }
      }
      __syncthreads();
      float dpart = 0.f;
      for (int j = tid; j < m; j += TD_THREADS) {
        float s2 = sp[j];
#pragma unroll
        for (int w = 0; w < 8; ++w) s2 += part[w * pa + j];
        s2 *= tau;
        sp[j] = s2;
        dpart += s2 * sv[j];
      }
      tick(3);
      const float dot = block_sum_1024(dpart, red);
      const float half = 0.5f * tau * dot;
      for (int j = tid; j < m; j += TD_THREADS) sw[j] = sp[j] - half * sv[j];
      __syncthreads();
      tick(4);
      // ---- A22 -= v w^T + w v^T (lower triangle), same row / lane ownership ----
      float wr[TD_NC];
#pragma unroll
      for (int c = 0; c < TD_NC; ++c) { const int j = 32 * c + lane; wr[c] = (j < m) ? sw[j] : 0.f; }
      for (int i = warp; i < m; i += TD_THREADS / 32) {
        float* row = Lp + tri(o + i) + o;
        const float vi = sv[i], wi = sw[i];
#pragma unroll
        for (int c = 0; c < TD_NC; ++c) {
          if (32 * c > i) break;
          const int j = 32 * c + lane;
          if (j <= i) row[j] -= vi * wr[c] + wi * vr[c];
        }
      }
      __syncthreads();
      tick(5);
    }
  }
  if (timed) for (int i = 0; i < 8; ++i) g_td_clk[i] = acc_clk[i];
  if (tid == 0) {
    for (int k = (p >= 2 ? p - 2 : 0); k < p; ++k) {
      dT[(int64_t)b * ldt + k] = Lp[tri(k) + k];
      eT[(int64_t)b * ldt + k] = (k + 1 < p) ? Lp[tri(k + 1) + k] : gmax_all;
      tau_out[(int64_t)b * ldt + k] = 0.f;
    }
  }
}

// Apply the reflectors of slice b to the columns of a small matrix held as M[i * s_i + f * s_f] (i < p rows, f < F columns), in place:
// forward (reverse = 0): M <- H_{p-3} ... H_0 M = Q^T M;  reverse = 1: M <- H_0 ... H_{p-3} M = Q M.  One warp owns a column (the
// steps of one column need no block barrier); grid (ceil(F / 8), B), block 256.
__global__ void __launch_bounds__(256) reflectors_small_kernel(const float* __restrict__ V, int ldv, const float* __restrict__ tau, int ldt, int p,
                                                               float* __restrict__ M, int64_t batch_stride, int64_t s_i, int64_t s_f, int F,
                                                               int reverse) {
  __shared__ float col[8][MAX_COLS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.y, f = blockIdx.x * 8 + warp;
  if (f >= F) return;
  float* Mb = M + (int64_t)b * batch_stride + (int64_t)f * s_f;
  float* c = col[warp];
  for (int i = lane; i < p; i += 32) c[i] = Mb[(int64_t)i * s_i];
  __syncwarp();
  const float* Vb = V + (int64_t)b * p * ldv;
  for (int kk = 0; kk + 2 < p; ++kk) {
    const int k = reverse ? (p - 3 - kk) : kk;
    const float t = tau[(int64_t)b * ldt + k];
    if (t == 0.f) continue;
    const float* v = Vb + (int64_t)k * ldv;
    float s = 0.f;
    for (int i = k + 1 + lane; i < p; i += 32) s += v[i] * c[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    s *= t;
    for (int i = k + 1 + lane; i < p; i += 32) c[i] -= s * v[i];
    __syncwarp();
  }
  for (int i = lane; i < p; i += 32) Mb[(int64_t)i * s_i] = c[i];
}

// L D L^T of T_b + alpha I for every (slice, alpha) and the reduced-basis solutions x = (T + alpha I)^-1 bt for the F columns of
// bt = Q^T X^T y ([B][Pp][F]).  One thread per (slice, alpha): the recurrences are sequential in k and tiny.
// lfac / dinv: [B][A][ldt];  xout: [B][A][F][Pp] (zero beyond p).  A pivot <= pivot_tol * (largest diagonal entry of G + alpha I) flags the slice.
__global__ void tri_ldl_solve_kernel(const float* __restrict__ dT, const float* __restrict__ eT, int ldt, int p, int Pp, const float* __restrict__ alphas,
                                     int A, int B, const float* __restrict__ bt, int F, float pivot_tol, float* __restrict__ lfac,
                                     float* __restrict__ dinv, float* __restrict__ xout, int* __restrict__ status) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * A) return;
  const int b = e / A, a = e - b * A;
  const float alpha = alphas[a];
  const float* d = dT + (int64_t)b * ldt;
  const float* ev = eT + (int64_t)b * ldt;
  float* l = lfac + (int64_t)e * ldt;
  float* di = dinv + (int64_t)e * ldt;
  const float floor_ = fmaxf((ev[p - 1] + alpha) * pivot_tol, 1e-30f);      // ev[p - 1] = largest diagonal entry of G_b (tridiag_kernel)
  bool bad = false;
  float dk = d[0] + alpha;
  for (int k = 0; k < p; ++k) {
    if (!(dk > floor_)) { bad = true; dk = floor_; }
    const float inv = 1.f / dk;
    di[k] = inv;
    if (k + 1 < p) {
      const float lk = ev[k] * inv;
      l[k] = lk;
      dk = d[k + 1] + alpha - lk * ev[k];
    } else {
      l[k] = 0.f;
    }
  }
  if (bad) atomicOr(&status[b], 1);
  for (int f = 0; f < F; ++f) {
    float* x = xout + ((int64_t)e * F + f) * Pp;
    const float* rhs = bt + (int64_t)b * Pp * F + f;
    float u = 0.f;
    for (int k = 0; k < p; ++k) {                 // forward: u_k = b_k - l_{k-1} u_{k-1};  y_k = u_k / d_k
      u = rhs[(int64_t)k * F] - (k > 0 ? l[k - 1] * u : 0.f);
      x[k] = u * di[k];
    }
    for (int k = p - 2; k >= 0; --k) x[k] -= l[k] * x[k + 1];      // backward
    for (int k = p; k < Pp; ++k) x[k] = 0.f;
  }
}

// grid (ceil(n / 32), B), block 256: a 32-row tile of Xc_b goes to shared memory, all reflectors are applied (Z = Xc Q, written back in
// place), and the leverages of the tile's rows for every alpha follow from the tridiagonal factors:  d_out[(b A + a) n + i] = 1 - h - c0.
// The reflector rows arrive through an 8-deep cp.async ring (first version: one global load per step = one L2 round trip per
// reflector, 71 ms per fold; the step itself is ~80 FMAs per thread).
constexpr int RA_ROWS = 32;
constexpr int RA_RING = 8;
inline size_t rows_apply_smem(int Pp, int p, int A) {
  return ((size_t)RA_ROWS * (Pp + 1) + (size_t)RA_RING * Pp + (size_t)((p + 3) & ~3) + 2 * (size_t)A * p + 8) * 4;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ long long g_ra_clk[4];       // developer timing: cycles of block (0, 0) in load / reflectors / write-back / leverages (tools/tri_phases.py)
__global__ void __launch_bounds__(256) rows_apply_leverage_kernel(float* __restrict__ Xc, int n, int p, int Pp, const float* __restrict__ V, int ldv,
                                                                  const float* __restrict__ tau, int ldt, const float* __restrict__ lfac,
                                                                  const float* __restrict__ dinv, int A, float c0, float* __restrict__ d_out,
                                                                  int apply) {
  // apply = 1: reflectors + write-back (+ leverages when A > 0);  apply = 0: the rows already are in the reduced basis, leverages only
  extern __shared__ __align__(16) float sm[];
  const int LDT = Pp + 1;
  float* ring = sm;                                        // [RA_RING][Pp], 16-byte aligned rows (Pp % 16 == 0)
  float* tile = ring + RA_RING * Pp;
  float* stau = tile + RA_ROWS * LDT + ((4 - ((RA_ROWS * LDT) & 3)) & 3);
  float* sl = stau + ((p + 3) & ~3);
  float* sd = sl + A * p;
  const int tid = threadIdx.x, b = blockIdx.y, row0 = blockIdx.x * RA_ROWS;
  const bool timed = blockIdx.x == 0 && blockIdx.y == 0 && tid == 0;
  long long t_last = timed ? clock64() : 0;
  auto tick = [&](int ph) { if (timed) { const long long t = clock64(); g_ra_clk[ph] = t - t_last; t_last = t; } };
  float* Xb = Xc + ((int64_t)b * n + row0) * Pp;
  const int nrows = min(RA_ROWS, n - row0);
  const float* Vb = V + (int64_t)b * p * ldv;
  const int nref = (apply && p >= 3) ? p - 2 : 0;          // reflectors 0 .. p - 3
  const int chunks = Pp / 4;                               // 16-byte chunks of a reflector row (ldv == Pp)
  auto issue = [&](int k) {                                // reflector k -> ring slot k % RA_RING (one commit group per reflector, possibly empty)
    if (k < nref && tid < chunks) cp_async16(ring + (k % RA_RING) * Pp + tid * 4, Vb + (int64_t)k * ldv + tid * 4);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int k = 0; k < RA_RING - 1; ++k) issue(k);
  for (int e = tid; e < RA_ROWS * Pp; e += 256) {
    const int r = e / Pp, c = e - r * Pp;
    tile[r * LDT + c] = (r < nrows) ? Xb[(int64_t)r * Pp + c] : 0.f;
  }
  for (int e = tid; e < A * p; e += 256) {
    const int a = e / p, k = e - a * p;
    sl[e] = lfac[((int64_t)b * A + a) * ldt + k];
    sd[e] = dinv[((int64_t)b * A + a) * ldt + k];
  }
  for (int k = tid; k < p; k += 256) stau[k] = tau[(int64_t)b * ldt + k];
  // eight lanes per row; lane q8 keeps the row's 16-byte chunks c = q8 (mod 8) in registers (Pp <= 320: at most 10 chunks), so a reflector
  // costs one 16-byte shared-memory load per four FMAs and no stores (with the tile in shared memory: two loads + a store per FMA pair,
  // measured 64 ms per fold)
  const int r = tid >> 3, q8 = tid & 7;
  constexpr int NCH = MAX_COLS / 32;                        // chunks per lane
  float4 reg[NCH];
  __syncthreads();                                          // the tile is complete
  tick(0);
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c4 = 4 * (q8 + 8 * i);
    reg[i] = (c4 < Pp) ? make_float4(tile[r * LDT + c4], tile[r * LDT + c4 + 1], tile[r * LDT + c4 + 2], tile[r * LDT + c4 + 3])
                       : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int k = 0; k < nref; ++k) {
    asm volatile("cp.async.wait_group %0;" ::"n"(RA_RING - 2) : "memory");      // reflector k has landed (this thread's part)
    __syncthreads();                                        // ... everybody's part; and step k - 1 is finished everywhere
    issue(k + RA_RING - 1);                                 // into the slot step k - 1 just released
    const float t = stau[k];
    if (t != 0.f) {
      const float4* sv4 = reinterpret_cast<const float4*>(ring + (k % RA_RING) * Pp);
      float4 v[NCH];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = q8 + 8 * i;                           // chunk index; v_k is zero up to column k, chunks entirely below are skipped
        if (4 * c + 3 > k && 4 * c < Pp) {
          v[i] = sv4[c];
          s += reg[i].x * v[i].x + reg[i].y * v[i].y + reg[i].z * v[i].z + reg[i].w * v[i].w;
        }
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s *= t;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = q8 + 8 * i;
        if (4 * c + 3 > k && 4 * c < Pp) {
          reg[i].x -= s * v[i].x; reg[i].y -= s * v[i].y; reg[i].z -= s * v[i].z; reg[i].w -= s * v[i].w;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NCH; ++i) {                           // rows back to the tile (leverage phase and write-back read it)
    const int c4 = 4 * (q8 + 8 * i);
    if (c4 < Pp) { tile[r * LDT + c4] = reg[i].x; tile[r * LDT + c4 + 1] = reg[i].y; tile[r * LDT + c4 + 2] = reg[i].z; tile[r * LDT + c4 + 3] = reg[i].w; }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  tick(1);
  if (apply)
    for (int e = tid; e < RA_ROWS * Pp; e += 256) {         // Z back in place
      const int rr = e / Pp, c = e - rr * Pp;
      if (rr < nrows) Xb[(int64_t)rr * Pp + c] = tile[rr * LDT + c];
    }
  tick(2);
  for (int pr = tid; pr < RA_ROWS * A; pr += 256) {         // leverages: lanes = rows (conflict-free tile reads), alpha uniform per warp
    const int rr = pr & (RA_ROWS - 1), a = pr / RA_ROWS;
    if (rr >= nrows) continue;
    const float* z = tile + rr * LDT;
    const float* l = sl + a * p;
    const float* di = sd + a * p;
    float w = 0.f, h = 0.f;
    for (int k = 0; k < p; ++k) {
      w = z[k] - (k > 0 ? l[k - 1] * w : 0.f);
      h += w * w * di[k];
    }
    d_out[((int64_t)b * A + a) * n + row0 + rr] = 1.f - h - c0;
  }
  tick(3);
}

// gather tables of the basis-change product: roffQ[b][k] = b Pp^2 + k (row k' of the B operand = column k' of Q_b), koffQ[j] = j Pp
__global__ void dense_q_tables_kernel(int B, int Pp, int64_t* __restrict__ roffQ, int64_t* __restrict__ koffQ) {
  const int64_t total = (int64_t)B * Pp;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    roffQ[e] = (e / Pp) * (int64_t)Pp * Pp + (e % Pp);
    if (e < Pp) koffQ[e] = e * (int64_t)Pp;
  }
}

// Q workspace: B identity matrices [Pp][Pp] (rows >= p stay zero rows with a unit diagonal: they are never read as real columns)
__global__ void identity_batched_kernel(float* __restrict__ Q, int B, int Pp) {
  const int64_t total = (int64_t)B * Pp * Pp;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % Pp), r = (int)((e / Pp) % Pp);
    Q[e] = (r == c) ? 1.f : 0.f;
  }
}

// intercept[b][f] = ybar - <coef[b][f], mu[b]>  (after the coefficients are back in the original basis); one warp per (b, f)
__global__ void intercept_batched_kernel(const float* __restrict__ coef, const float* __restrict__ mu, const float* __restrict__ ybar, int B, int F,
                                         int p, int Pp, int fit_intercept, float* __restrict__ intercept) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B * F) return;
  const int b = w / F;
  double m = 0;
  for (int c = lane; c < p; c += 32) m += (double)coef[(int64_t)w * p + c] * (double)mu[(int64_t)b * Pp + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
  if (lane == 0) intercept[w] = fit_intercept ? (float)((double)ybar[w] - m) : 0.f;
}

// compact copy of the centred Grams: out[b][p][p] <- G[b][Pp][Pp]
__global__ void copy_gram_kernel(const float* __restrict__ G, int B, int p, int Pp, float* __restrict__ out) {
  const int64_t total = (int64_t)B * p * p;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % p);
    const int64_t t = e / p;
    const int r = (int)(t % p), b = (int)(t / p);
    out[e] = G[((int64_t)b * Pp + r) * Pp + c];
  }
}

// yhat_b[i][f] = sum_c X_b[i][c] coef[b][f][c] + intercept[b][f], written through an output view (sample / target / batch strides)
// grid (ceil(B / 32), n, ceil(F / 4)), block (32, 4): threads over slices (contiguous for a Sliding view), 4 samples per block row
__global__ void predict_kernel(View x, int n, int p, int B, const float* __restrict__ coef, const float* __restrict__ icpt, int F,
                               float* __restrict__ out, int64_t o_sample, int64_t o_target, int64_t o_batch) {
  const int b = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y, f0 = blockIdx.z * 4;
  if (b >= B || i >= n) return;
  const int nf = min(4, F - f0);
  const float* src = x.base + (int64_t)i * x.s_sample + (int64_t)b * x.s_batch;
  const float* w = coef + ((int64_t)b * F + f0) * p;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = 0; c < p; ++c) {
    const float xv = __ldg(src + (int64_t)c * x.s_col);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nf) acc[j] += xv * __ldg(w + (int64_t)j * p + c);
  }
  for (int j = 0; j < nf; ++j)
    out[(int64_t)i * o_sample + (int64_t)(f0 + j) * o_target + (int64_t)b * o_batch] = acc[j] + (icpt ? icpt[(int64_t)b * F + f0 + j] : 0.f);
}

}  // namespace db
}  // namespace mvb
