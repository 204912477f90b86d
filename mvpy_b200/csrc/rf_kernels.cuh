// Kernels behind the ReceptiveField estimator (mvpy/estimators/receptivefield.py:779-1556).
// The heavy part -- per-epoch lagged Gram and cross-moment -- is the implicit-design contraction kernel run on a
// zero-padded stimulus; what is here turns its output into the reference's cov_ layout, sums epochs over a fold and
// solves the (non-symmetric, see assemble_kernel) penalised normal equations for a batch of penalties.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace mvb {
namespace rf {

// cov[n][(i,a)][(j,b)] in the reference's order (feature-major, lag k = a ascending) from the Gram G[n] of the
// zero-padded design in kernel order (lag index w = W-1-k).  Two reference behaviours are reproduced on purpose:
//  * blocks below the diagonal (i > j) are a plain copy of block (j, i), not its transpose (receptivefield.py:1097-1103);
//  * with edge correction and negative lags, the term subtracted for rows before the epoch start is
//    q[a][b] = sum_{d<=min(a,b)} x_i[L-1-a+d] x_j[L-1-b+d]  (L = -s_min; :801-807 on the flipped epoch) where the rows
//    that really lie before the start give  e[a][b] = sum_{d<min(L-a,L-b)} x_i[L-a-1-d] x_j[L-b-1-d];  G already
//    excludes e (its rows are 0..T-1), so the reference value is G + e - q on the leading Lh x Lh corner.
template <typename T>
__global__ void assemble_kernel(const T* __restrict__ G, const T* __restrict__ X, int N, int C, int W, int Tn, int L, int Lh,
                                T* __restrict__ cov) {
  const int64_t p = (int64_t)C * W;
  const int64_t total = (int64_t)N * p * p;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = e / (p * p);
    const int64_t rc = e - n * p * p;
    const int r = (int)(rc / p), c = (int)(rc - (int64_t)r * p);
    int i = r / W, j = c / W;
    const int a = r - i * W, b = c - j * W;
    if (i > j) { const int t = i; i = j; j = t; }
    double v = (double)G[n * p * p + ((int64_t)i * W + (W - 1 - a)) * p + ((int64_t)j * W + (W - 1 - b))];
    if (a < Lh && b < Lh) {
      const T* xi = X + ((int64_t)n * C + i) * Tn;
      const T* xj = X + ((int64_t)n * C + j) * Tn;
      const int nq = (a < b ? a : b) + 1, ne = L - (a > b ? a : b);
      double q = 0.0, ex = 0.0;
      for (int d = 0; d < nq; ++d) q += (double)xi[L - 1 - a + d] * (double)xj[L - 1 - b + d];
      for (int d = 0; d < ne; ++d) ex += (double)xi[L - a - 1 - d] * (double)xj[L - b - 1 - d];
      v += ex - q;
    }
    cov[e] = (T)v;
  }
}

// out[e] = sum_i A[idx[i]][e]  (epochs of a fold, receptivefield.py:1204-1206)
template <typename T>
__global__ void sum_slabs_kernel(const T* __restrict__ A, int64_t slab, const int32_t* __restrict__ idx, int n_idx,
                                 T* __restrict__ out) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < slab; e += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int i = 0; i < n_idx; ++i) s += (double)A[(int64_t)idx[i] * slab + e];
    out[e] = (T)s;
  }
}

// mats[s] = G + alpha[s] * reg   (receptivefield.py:1224-1226), one system per penalty
template <typename T>
__global__ void penalised_kernel(const T* __restrict__ G, const T* __restrict__ reg, const T* __restrict__ alphas, int S,
                                 int64_t pp, T* __restrict__ mats) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < pp * S; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = e / pp, k = e - s * pp;
    mats[e] = G[k] + alphas[s] * reg[k];
  }
}

// Batched dense solve A x = B by LU with partial pivoting (the reference calls torch.linalg.solve, :1229, on a matrix
// that is not symmetric).  One CTA per system; A (p x p, row-major) and B (p x F) are overwritten, B with the solution.
// info[s] = 0, or k+1 when column k has no non-zero pivot.
constexpr int LU_THREADS = 1024;
template <typename T>
__global__ void __launch_bounds__(LU_THREADS) lu_solve_kernel(T* __restrict__ A_all, T* __restrict__ B_all, int p, int F, int* info) {
  T* A = A_all + (int64_t)blockIdx.x * p * p;
  T* B = B_all + (int64_t)blockIdx.x * p * F;
  __shared__ double s_val[LU_THREADS / 32];
  __shared__ int s_idx[LU_THREADS / 32];
  __shared__ int s_piv;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  for (int k = 0; k < p; ++k) {
    double best = -1.0; int bi = k;
    for (int r = k + tid; r < p; r += nt) { const double v = fabs((double)A[(int64_t)r * p + k]); if (v > best) { best = v; bi = r; } }
    for (int o = 16; o; o >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, best, o); const int oi = __shfl_down_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      best = lane < nt / 32 ? s_val[lane] : -1.0; bi = lane < nt / 32 ? s_idx[lane] : p;
      for (int o = 16; o; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, best, o); const int oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (lane == 0) s_piv = best > 0.0 ? bi : -1;
    }
    __syncthreads();
    const int piv = s_piv;
    if (piv < 0) { if (tid == 0) info[blockIdx.x] = k + 1; return; }
    if (piv != k) {
      for (int c = k + tid; c < p + F; c += nt) {
        T* a = c < p ? &A[(int64_t)k * p + c] : &B[(int64_t)k * F + (c - p)];
        T* b = c < p ? &A[(int64_t)piv * p + c] : &B[(int64_t)piv * F + (c - p)];
        const T t = *a; *a = *b; *b = t;
      }
      __syncthreads();
    }
    const T inv = (T)1 / A[(int64_t)k * p + k];
    const int rem = p - 1 - k, wide = rem + F;             // trailing columns of A, then the right-hand sides
    for (int64_t e = tid; e < (int64_t)rem * wide; e += nt) {
      const int r = k + 1 + (int)(e / wide), cc = (int)(e % wide);
      const T l = A[(int64_t)r * p + k] * inv;
      if (cc < rem) A[(int64_t)r * p + k + 1 + cc] -= l * A[(int64_t)k * p + k + 1 + cc];
      else B[(int64_t)r * F + (cc - rem)] -= l * B[(int64_t)k * F + (cc - rem)];
    }
    __syncthreads();
  }
  for (int k = p - 1; k >= 0; --k) {
    const T inv = (T)1 / A[(int64_t)k * p + k];
    for (int f = tid; f < F; f += nt) B[(int64_t)k * F + f] *= inv;
    __syncthreads();
    for (int64_t e = tid; e < (int64_t)k * F; e += nt) {
      const int r = (int)(e / F), f = (int)(e % F);
      B[(int64_t)r * F + f] -= A[(int64_t)r * p + k] * B[(int64_t)k * F + f];
    }
    __syncthreads();
  }
  if (tid == 0) info[blockIdx.x] = 0;
}

}  // namespace rf

// ---- classification metrics, same [R][K] layout as score_kernel: reduced samples r are the slow axis ----
// acc[k] = mean_r [ y[r][k] == yh[r][k] ]                                   (math/accuracy.py:13-33)
template <typename T>
__global__ void accuracy_kernel(const T* __restrict__ y, const T* __restrict__ yh, int64_t R, int64_t K, T* __restrict__ acc) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  int64_t hit = 0;
  for (int64_t r = 0; r < R; ++r) hit += y[r * K + k] == yh[r * K + k];
  acc[k] = (T)((double)hit / (double)R);
}

// auc[k] = (sum of the average ranks of the positives - P (P + 1) / 2) / (P N)   (math/roc_auc.py:139-152), written as the
// pair count  sum_{i pos} sum_{j neg} [s_i > s_j] + [s_i == s_j] / 2,  which is the same number; NaN without both classes.
template <typename T>
__global__ void auc_kernel(const T* __restrict__ pos, const T* __restrict__ score, int64_t R, int64_t K, T* __restrict__ auc) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double wins = 0.0;
  int64_t P = 0;
  for (int64_t i = 0; i < R; ++i) {
    if (pos[i * K + k] == (T)0) continue;
    ++P;
    const T si = score[i * K + k];
    int64_t gt = 0, eq = 0;
    for (int64_t j = 0; j < R; ++j) {
      if (pos[j * K + k] != (T)0) continue;
      const T sj = score[j * K + k];
      gt += si > sj; eq += si == sj;
    }
    wins += (double)gt + 0.5 * (double)eq;
  }
  const int64_t N = R - P;
  auc[k] = (P == 0 || N == 0) ? (T)NAN : (T)(wins / ((double)P * (double)N));
}

}  // namespace mvb
