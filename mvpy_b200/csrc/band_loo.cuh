// band_loo.cuh -- (T + alpha I)^-1 for a block-tridiagonal T (128 x 128 blocks) and a whole alpha grid.
//
// With Gp = Qt^T T Qt (lanczos_band.cuh), the LOO quantities of mvpy/estimators/ridgecv.py:230-251 become
//   leverage  h_a,i = || L_a^-1 z_i ||^2,            z_i = Qt (x_i - mu)         (T + aI = L_a L_a^T)
//   beta_a        = Qt^T (T + aI)^-1 Qt b
// L_a is block lower-bidiagonal:  C_j = A_j + aI - W_j W_j^T,  W_j = B_j L_{j-1}^-T,  L_j = chol(C_j).
// block_chol_kernel builds, per alpha (one CTA each, all in shared memory), the three matrices the solves need:
//   Minv_j = L_j^-1,   N_j = L_j^-1 W_j,   Pb_j = W_{j+1} L_j^-1
// so that forward substitution is  u_j = Minv_j z_j - N_j u_{j-1}  and back substitution
// x_j = Minv_j^T u_j - Pb_j^T x_{j+1}; both run as batched (over alpha) tensor-core contractions over all rows.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tc_gemm.cuh"
#include "jacobi_block.cuh"

namespace mvb {
namespace bl {

constexpr int NB = 128;
constexpr int LD = NB + 1;
constexpr int CHOL_SMEM = 3 * NB * LD * 4;

// out[r][c] (global, ld NB) = sum_k X[r][k] * Y[k][c] for shared-memory X, Y (ld LD); 1024 threads, 4x4 per thread.
// tri: 0 full; 1: X lower-triangular (k <= r); 2: Y lower-triangular (k >= c)
__device__ __forceinline__ void smem_matmul(const float* X, const float* Y, float* out_g, float* out_s, int tri, float sign,
                                            const float* add_s, float* neg_g = nullptr, int neg_ld = 0) {
  const int tr = (threadIdx.x >> 5) * 4, tcn = (threadIdx.x & 31) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  int k0 = 0, k1 = NB;
  if (tri == 1) k1 = tr + 4;
  if (tri == 2) k0 = tcn;
  for (int k = k0; k < k1; ++k) {
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = X[(tr + i) * LD + k];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = Y[k * LD + tcn + j];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = sign * acc[i][j];
      if (add_s) v += add_s[(tr + i) * LD + tcn + j];
      if (out_g) out_g[(tr + i) * NB + tcn + j] = v;
      if (out_s) out_s[(tr + i) * LD + tcn + j] = v;
      if (neg_g) neg_g[(tr + i) * neg_ld + tcn + j] = -v;
    }
}

// one CTA per alpha
__global__ void __launch_bounds__(1024) block_chol_kernel(const float* __restrict__ Tdiag, const float* __restrict__ Tsub, int nblk,
                                                          const float* __restrict__ alphas, float* Minv, float* Nmat, float* Pb,
                                                          float* MN /* [A][nblk][128][256]: row c = [Minv_j[c][:], -N_j[c][:]] */) {
  extern __shared__ float sm[];
  float* Cm = sm;                  // A_j + aI - W W^T  -> L_j -> (scratch)
  float* Wm = sm + NB * LD;        // W_j
  float* Li = sm + 2 * NB * LD;    // L_{j-1}^-1, then L_j^-1
  const int a = blockIdx.x;
  const float alpha = alphas[a];
  // Pivot floor (modified Cholesky): T is the computed image of a PSD Gram, so for a (numerically) singular design its
  // smallest eigenvalues are rounding noise of either sign, ~ eps * max diag(T).  Pivots below 8 eps * max diag are
  // raised to it: the deficient directions get a finite, tiny weight instead of NaN -- what the reference's SVD does
  // with its noise-level singular values (ridgecv.py:223-233) -- and well-conditioned designs never reach the floor.
  __shared__ float s_tmax;
  {
    __shared__ float red[32];
    float m = 0.f;
    for (int e = threadIdx.x; e < nblk * NB; e += blockDim.x) m = fmaxf(m, fabsf(Tdiag[(int64_t)(e / NB) * NB * NB + (e % NB) * (NB + 1)]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t = fmaxf(t, red[w]); s_tmax = t; }
    __syncthreads();
  }
  const float pivot_floor = fmaxf(4.8e-7f * s_tmax, 1e-30f);
  float* Minv_a = Minv + (int64_t)a * nblk * NB * NB;
  float* N_a = Nmat + (int64_t)a * nblk * NB * NB;
  float* Pb_a = Pb + (int64_t)a * nblk * NB * NB;
  for (int j = 0; j < nblk; ++j) {
    const float* A = Tdiag + (int64_t)j * NB * NB;
    if (j > 0) {
      // W = B_j Li^T : stage B_j in Cm, transpose Li into ... (Li lower-tri: W[r][c] = sum_{k<=c} B[r][k] Li[c][k])
      const float* B = Tsub + (int64_t)j * NB * NB;
      for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) Cm[(e / NB) * LD + (e % NB)] = B[e];
      __syncthreads();
      {
        const int tr = (threadIdx.x >> 5) * 4, tcn = (threadIdx.x & 31) * 4;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.f;
        for (int k = 0; k < tcn + 4; ++k) {
          float av[4], bv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) av[i] = Cm[(tr + i) * LD + k];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) bv[jj] = (k <= tcn + jj) ? Li[(tcn + jj) * LD + k] : 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[i][jj] += av[i] * bv[jj];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) Wm[(tr + i) * LD + tcn + jj] = acc[i][jj];
      }
      __syncthreads();
      // Pb_{j-1} = W Li  (Li lower-triangular)
      smem_matmul(Wm, Li, Pb_a + (int64_t)(j - 1) * NB * NB, nullptr, 2, 1.f, nullptr);
      __syncthreads();
      // Cm = A_j + aI - W W^T   (needs W^T: out[r][c] = sum_k W[r][k] W[c][k])
      {
        const int tr = (threadIdx.x >> 5) * 4, tcn = (threadIdx.x & 31) * 4;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.f;
        for (int k = 0; k < NB; ++k) {
          float av[4], bv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) av[i] = Wm[(tr + i) * LD + k];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) bv[jj] = Wm[(tcn + jj) * LD + k];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[i][jj] += av[i] * bv[jj];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int r = tr + i, c = tcn + jj;
            Cm[r * LD + c] = A[r * NB + c] + (r == c ? alpha : 0.f) - acc[i][jj];
          }
      }
    } else {
      for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
        const int r = e / NB, c = e % NB;
        Cm[r * LD + c] = A[e] + (r == c ? alpha : 0.f);
      }
    }
    __syncthreads();
    // in-place modified Cholesky (lower) of Cm: pivot_k = max(C_kk, floor, max_i C_ik^2 / tmax) -- the third term
    // (Gill-Murray) bounds |L_ik| by sqrt(tmax) when rounding made the matrix indefinite, so the factor cannot blow up.
    // 16-column panels: (1) all threads bring the panel up to date with the columns already factored (left-looking),
    // (2) threads 0..127 (one matrix row each, its 16 panel entries in registers) factor the panel column by column
    // with two 128-thread named barriers per column instead of five block barriers.
    {
      constexpr int PB = 16;
      __shared__ float s_red[4], s_diag, s_lcol[PB];
      const float inv_tmax = 1.f / fmaxf(s_tmax, 1e-30f);
      for (int pb = 0; pb < NB; pb += PB) {
        for (int e = threadIdx.x; e < (NB - pb) * PB; e += blockDim.x) {
          const int i = pb + e / PB, c = pb + e % PB;
          if (c <= i) {
            float acc = 0.f;
            for (int k = 0; k < pb; ++k) acc += Cm[i * LD + k] * Cm[c * LD + k];
            Cm[i * LD + c] -= acc;
          }
        }
        __syncthreads();
        if (threadIdx.x < NB) {
          const int r = threadIdx.x;
          float x[PB];
#pragma unroll
          for (int c = 0; c < PB; ++c) x[c] = (r >= pb + c) ? Cm[r * LD + pb + c] : 0.f;
#pragma unroll
          for (int k = 0; k < PB; ++k) {
            const int col = pb + k;
            float v = (r > col) ? fabsf(x[k]) : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
            if ((r & 31) == 0) s_red[r >> 5] = v;
            if (r == col) s_diag = x[k];
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const float cm = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
            const float d = sqrtf(fmaxf(fmaxf(s_diag, pivot_floor), cm * cm * inv_tmax));
            if (r == col) x[k] = d;
            else if (r > col) x[k] /= d;
            if (r > col && r < pb + PB) s_lcol[r - pb] = x[k];
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (r > col) {
#pragma unroll
              for (int c = 0; c < PB; ++c)
                if (c > k && pb + c <= r) x[c] -= x[k] * s_lcol[c];
            }
          }
#pragma unroll
          for (int c = 0; c < PB; ++c)
            if (r >= pb + c) Cm[r * LD + pb + c] = x[c];
        }
        __syncthreads();
      }
    }
    // Li = L^-1 (lower) by 16 x 16 blocks:  M_ii = L_ii^-1 (all eight diagonal blocks at once, one thread per column),
    // then block row by block row  M_ij = -M_ii sum_{j <= k < i} L_ik M_kj  (the strictly-lower blocks of Cm become scratch).
    {
      constexpr int PB = 16, NBLK = NB / PB;
      for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) Li[(e / NB) * LD + (e % NB)] = 0.f;
      __syncthreads();
      if (threadIdx.x < NB) {
        const int blk = threadIdx.x / PB, c = threadIdx.x % PB, o = blk * PB;
        float x[PB];
#pragma unroll
        for (int i = 0; i < PB; ++i) {
          float v = (i == c) ? 1.f : 0.f;
#pragma unroll
          for (int k = 0; k < PB; ++k)
            if (k < i && k >= c) v -= Cm[(o + i) * LD + o + k] * x[k];
          x[i] = (i < c) ? 0.f : v / Cm[(o + i) * LD + o + i];
        }
#pragma unroll
        for (int i = 0; i < PB; ++i) Li[(o + i) * LD + o + c] = x[i];
      }
      __syncthreads();
      for (int bi = 1; bi < NBLK; ++bi) {
        float t_acc[2] = {0.f, 0.f};
        int n_mine = 0;
        for (int e = threadIdx.x; e < bi * PB * PB; e += blockDim.x, ++n_mine) {
          const int jb = e / (PB * PB), r = (e / PB) % PB, c = e % PB;
          float acc = 0.f;
          for (int k = jb * PB; k < bi * PB; ++k) acc += Cm[(bi * PB + r) * LD + k] * Li[k * LD + jb * PB + c];
          t_acc[n_mine] = acc;
        }
        __syncthreads();
        n_mine = 0;
        for (int e = threadIdx.x; e < bi * PB * PB; e += blockDim.x, ++n_mine) {
          const int jb = e / (PB * PB), r = (e / PB) % PB, c = e % PB;
          Cm[(bi * PB + r) * LD + jb * PB + c] = t_acc[n_mine];
        }
        __syncthreads();
        for (int e = threadIdx.x; e < bi * PB * PB; e += blockDim.x) {
          const int jb = e / (PB * PB), r = (e / PB) % PB, c = e % PB;
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < PB; ++k) acc += Li[(bi * PB + r) * LD + bi * PB + k] * Cm[(bi * PB + k) * LD + jb * PB + c];
          Li[(bi * PB + r) * LD + jb * PB + c] = -acc;
        }
        __syncthreads();
      }
    }
    // outputs: Minv_j = Li ; N_j = Li W (Li lower-triangular)
    float* MN_aj = MN + ((int64_t)a * nblk + j) * NB * 2 * NB;
    for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
      const float v = Li[(e / NB) * LD + (e % NB)];
      Minv_a[(int64_t)j * NB * NB + e] = v;
      MN_aj[(e / NB) * 2 * NB + (e % NB)] = v;
      if (j == 0) MN_aj[(e / NB) * 2 * NB + NB + (e % NB)] = 0.f;
    }
    // N_j = Li W, plus its negated copy for the fused forward substitution  u_j = [Minv_j | -N_j] [z_j ; u_{j-1}]
    if (j > 0) smem_matmul(Li, Wm, N_a + (int64_t)j * NB * NB, nullptr, 1, 1.f, nullptr, MN_aj + NB, 2 * NB);
    __syncthreads();
  }
}

// table[e] = e * 128  (row offsets of stacked 128-wide matrices)
__global__ void tab128_kernel(int64_t* t, int64_t n) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) t[e] = e * NB;
}
__global__ void tab_ld_kernel(int64_t* t, int64_t n, int64_t ld) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) t[e] = e * ld;
}

// U[row][0..128) (ld ldu, column offset applied by the caller) = C1 - C2 ; optional h[row] += sum U^2
__global__ void recur_update_kernel(const float* __restrict__ C1, const float* __restrict__ C2, int64_t rows, float* U, int64_t ldu,
                                    float* h) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float4 a = reinterpret_cast<const float4*>(C1 + r * NB)[lane];
  float4 u = a;
  if (C2) {
    const float4 b = reinterpret_cast<const float4*>(C2 + r * NB)[lane];
    u.x -= b.x; u.y -= b.y; u.z -= b.z; u.w -= b.w;
  }
  reinterpret_cast<float4*>(U + r * ldu)[lane] = u;
  if (h) {
    float s = u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w;
    s = bj::wsum(s);
    if (lane == 0) h[r] += s;
  }
}

}  // namespace bl
}  // namespace mvb
