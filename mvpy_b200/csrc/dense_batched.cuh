// dense_batched.cuh -- B independent dense leave-one-out ridge problems of one shape in a handful of launches.
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
      m += (double)v * (double)mu[(int64_t)b * Pp + c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
    if (lane == 0) intercept[(int64_t)b * F + f] = fit_intercept ? (float)((double)ybar[(int64_t)b * F + f] - m) : 0.f;
  }
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
