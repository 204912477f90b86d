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
// Shared memory: packed lower triangle (row i at tri(i)), one [Pp][16] panel, one 16 x 17 diagonal-block inverse.
inline size_t chol_inv_smem(int Pp) { return ((size_t)((tri(Pp) + 3) & ~3) + (size_t)Pp * NB + NB * (NB + 1)) * 4; }

__device__ __forceinline__ void diag_inverse(const float* __restrict__ Lp, int J, float* __restrict__ Dinv, int lane) {
  // inverse of the lower-triangular 16 x 16 diagonal block at J: lane c < 16 solves column c by forward substitution
  if (lane < NB) {
    const int c = lane;
    float x[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      float v = (i == c) ? 1.f : 0.f;
#pragma unroll
      for (int k = 0; k < NB; ++k)
        if (k < i && k >= c) v -= Lp[tri(J + i) + J + k] * x[k];
      x[i] = (i < c) ? 0.f : v / Lp[tri(J + i) + J + i];
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) Dinv[i * (NB + 1) + c] = x[i];
  }
}

__global__ void __launch_bounds__(CI_THREADS, 1) chol_inv_kernel(const float* __restrict__ G, int p, int Pp, const float* __restrict__ alphas,
                                                                int A, float pivot_tol, float* __restrict__ Li, int tri_rows,
                                                                int* __restrict__ status, long long* __restrict__ phase_clk = nullptr) {
  extern __shared__ __align__(16) float sm[];
  float* Lp = sm;
  float* panel = sm + ((tri(Pp) + 3) & ~3);
  float* Dinv = panel + Pp * NB;
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
  for (int e = tid; e < Pp * Pp; e += CI_THREADS) {
    const int i = e / Pp, j = e - i * Pp;
    if (j <= i) {
      float v = (i < p) ? Gb[e] : 0.f;
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
    if (warp == 0) {                               // (a) the 16 x 16 diagonal block: lane r owns row kb + r
      const int r = lane;
      bool bad = false;
      for (int k = 0; k < NB; ++k) {
        float piv = Lp[tri(kb + k) + kb + k];
        if (kb + k < p && !(piv > floor_)) { bad = true; piv = floor_; }      // (padding rows are an identity block: pivot 1)
        const float d = sqrtf(piv);
        __syncwarp();
        if (r == k) Lp[tri(kb + k) + kb + k] = d;
        if (r > k && r < NB) Lp[tri(kb + r) + kb + k] /= d;
        __syncwarp();
        if (r > k && r < NB) {
          const float lrk = Lp[tri(kb + r) + kb + k];
          for (int c = k + 1; c <= r; ++c) Lp[tri(kb + r) + kb + c] -= lrk * Lp[tri(kb + c) + kb + k];
        }
        __syncwarp();
      }
      if (bad && lane == 0) s_bad = 1;
    }
    __syncthreads();
    tick(1);
    for (int i = kb + NB + tid; i < Pp; i += CI_THREADS) {      // (b) rows below: L[i, panel] = A[i, panel] L_kk^-T
      float* row = Lp + tri(i) + kb;
      float x[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        float v = row[c];
#pragma unroll
        for (int k = 0; k < NB; ++k)
          if (k < c) v -= x[k] * Lp[tri(kb + c) + kb + k];
        x[c] = v / Lp[tri(kb + c) + kb + c];
      }
#pragma unroll
      for (int c = 0; c < NB; ++c) row[c] = x[c];
    }
    __syncthreads();
    tick(2);
    {                                                          // (c) trailing update in 4 x 4 tiles of the lower triangle
      const int base = kb + NB, nt = (Pp - base) / 4, ntiles = tri(nt);
      for (int e = tid; e < ntiles; e += CI_THREADS) {
        int ti = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
        while (tri(ti + 1) <= e) ++ti;
        while (tri(ti) > e) --ti;
        const int tj = e - tri(ti);
        const int i0 = base + 4 * ti, j0 = base + 4 * tj;
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
        const float* ra[4];
        const float* rb[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) { ra[r] = Lp + tri(i0 + r) + kb; rb[r] = Lp + tri(j0 + r) + kb; }
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          float av[4], bv[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) { av[r] = ra[r][k]; bv[r] = rb[r][k]; }
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] += av[r] * bv[c];
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (j0 + c <= i0 + r) Lp[tri(i0 + r) + j0 + c] -= acc[r][c];
      }
    }
    __syncthreads();
    tick(3);
  }

  // ---- in-place inverse of the triangular factor, block columns from the right ----
  for (int J = Pp - NB; J >= 0; J -= NB) {
    if (warp == 0) diag_inverse(Lp, J, Dinv, lane);
    __syncthreads();
    tick(4);
    for (int i = J + NB + tid; i < Pp; i += CI_THREADS) {       // panel[i] = L[i, J..J+16) Dinv
      const float* row = Lp + tri(i) + J;
      float l[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) l[k] = row[k];
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < NB; ++k)
          if (k >= c) v += l[k] * Dinv[k * (NB + 1) + c];
        panel[i * NB + c] = v;
      }
    }
    if (tid < NB * NB) {                                         // the diagonal block itself
      const int r = tid / NB, c = tid % NB;
      if (c <= r) Lp[tri(J + r) + J + c] = Dinv[r * (NB + 1) + c];
    }
    __syncthreads();
    tick(5);
    {                                                            // Inv[R, J] = -Inv[R, R] panel[R]   (R = rows below the block)
      const int base = J + NB, nrt = (Pp - base) / 4, ntiles = nrt * 4;
      for (int e = tid; e < ntiles; e += CI_THREADS) {
        // long rows first: tile e -> row group from the bottom, so that the threads of a warp have similar trip counts
        const int ti = nrt - 1 - e / 4, c0 = (e & 3) * 4;
        const int i0 = base + 4 * ti;
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
        const float* ra[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) ra[r] = Lp + tri(i0 + r);
        for (int k = base; k < i0; ++k) {                        // full part: k below the tile's first row
          const float4 t = *reinterpret_cast<const float4*>(panel + k * NB + c0);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float av = ra[r][k];
            acc[r][0] += av * t.x; acc[r][1] += av * t.y; acc[r][2] += av * t.z; acc[r][3] += av * t.w;
          }
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {                         // triangular corner
          const int k = i0 + kk;
          const float4 t = *reinterpret_cast<const float4*>(panel + k * NB + c0);
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if (kk <= r) {
              const float av = ra[r][k];
              acc[r][0] += av * t.x; acc[r][1] += av * t.y; acc[r][2] += av * t.z; acc[r][3] += av * t.w;
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) Lp[tri(i0 + r) + J + c0 + c] = -acc[r][c];
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
