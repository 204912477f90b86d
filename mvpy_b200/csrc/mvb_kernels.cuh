// mvb_kernels.cuh -- CUDA-core kernels of the ridge hot path (everything that is not the
// tcgen05 contraction core): scaler statistics, edge padding, offset tables, the fp64/validation
// contraction, centring, the Jacobi eigensolver, leverage / LOO-loss reductions, alpha selection,
// prediction layout and the fused R2 / Pearson scoring kernel.  Templated on float / double.
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <math.h>

namespace mvb {
namespace cg = cooperative_groups;

template <typename T> struct Acc { using type = double; };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// Scaler (preprocessing/scaler.py:355-360, 385-393).  X viewed as [R][K]; one thread per kept cell,
// consecutive threads on consecutive cells (coalesced); two passes with fp64 accumulators, which is
// how torch's CPU var kernel accumulates fp32 input (results are rounded to T at the end, so tiny
// variances underflow to 0 exactly like the reference -- SURVEY.md H5).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void scaler_fit_kernel(const T* __restrict__ X, int64_t R, int64_t K, T* mean, T* var, T* scale) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double s = 0;
  for (int64_t r = 0; r < R; ++r) s += (double)X[r * K + k];
  const double m = s / (double)R;
  const T mT = (T)m;
  double q = 0;
  for (int64_t r = 0; r < R; ++r) {
    const double d = (double)X[r * K + k] - m;
    q += d * d;
  }
  const T v = (T)(q / (double)(R - 1));  // R == 1 -> NaN like torch.var
  T sc = (T)sqrt((double)v);
  if (sc == (T)0 || sc != sc) sc = (T)1;
  mean[k] = mT;
  var[k] = v;
  scale[k] = sc;
}

template <typename T>
__global__ void scaler_apply_kernel(const T* __restrict__ X, int64_t total, int64_t K, const T* __restrict__ mean,
                                    const T* __restrict__ scale, int with_mean, int with_std, int inverse, T* out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = i % K;
    T x = X[i];
    if (!inverse) {
      if (with_mean) x = x - mean[k];
      if (with_std) x = x / scale[k];
    } else {
      if (with_std) x = x * scale[k];
      if (with_mean) x = x + mean[k];
    }
    out[i] = x;
  }
}

// (N,C,T) -> (N,C,T+left+right), replicated edges  (timedelayed.py:415-417: clamp(t + lag, 0, T-1))
template <typename T>
__global__ void pad_kernel(const T* __restrict__ X, int64_t rows, int Tn, int left, int Tp, T* out) {
  const int64_t total = rows * Tp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / Tp;
    int j = (int)(i - r * Tp) - left;
    j = j < 0 ? 0 : (j > Tn - 1 ? Tn - 1 : j);
    out[i] = X[r * Tn + j];
  }
}

// Per-feature pivot of the padded stimulus: shift[c] = mean over the selected trials and all Tp samples of feature c (fp64
// accumulation), one block per feature.  Centring "before multiplying" (ridgecv.py:59-94 centres the design itself): the
// centred Gram / cross-moment do not change when a constant is subtracted from a column, and every column (c, w) of the
// lag-expanded design has a mean within edge effects of its feature's mean, so subtracting the pivot from the stimulus removes
// the n mu mu^T cancellation from the rank-1 centring correction of G for inputs whose mean dwarfs their spread.
template <typename T>
__global__ void feature_pivot_kernel(const T* __restrict__ xp, const int32_t* __restrict__ trial_idx, int n_trials, int64_t trial_stride,
                                     int64_t feat_stride, int Tp, T* __restrict__ shift) {
  __shared__ double red[256];
  const int c = blockIdx.x;
  double s = 0.0;
  const int64_t total = (int64_t)n_trials * Tp;
  for (int64_t e = threadIdx.x; e < total; e += blockDim.x) {
    const int i = (int)(e / Tp), j = (int)(e - (int64_t)i * Tp);
    s += (double)xp[(int64_t)(trial_idx ? trial_idx[i] : i) * trial_stride + (int64_t)c * feat_stride + j];
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) shift[c] = (T)(red[0] / (double)total);
}
// out[n][c][j] = xp[n][c][j] - shift[c]   (all trials of the array, contiguous (N, C, Tp))
template <typename T>
__global__ void feature_shift_kernel(const T* __restrict__ xp, int64_t total, int C, int Tp, const T* __restrict__ shift, T* __restrict__ out) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = xp[e] - shift[(e / Tp) % C];
}

template <typename T>
__global__ void transpose_kernel(const T* __restrict__ X, int64_t R, int64_t C, T* out) {
  __shared__ T tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t r = r0 + j, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[j][threadIdx.x] = X[r * C + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < C) out[c * R + r] = tile[threadIdx.x][j];
  }
}

// ------------------------------------------------------------------------------------------------
// offset tables of the implicit design (rows i=(s,t), columns j=(f,w)) and of the targets
// ------------------------------------------------------------------------------------------------
__global__ void design_offsets_kernel(const int32_t* trial_idx, int n_trials, int n_time, int64_t trial_stride,
                                      const int32_t* feat_idx, int n_feat, int n_lag, int64_t feat_stride,
                                      int64_t y_trial_stride, int64_t* row_off, int64_t* col_off, int64_t* yrow_off) {
  const int64_t n = (int64_t)n_trials * n_time, p = (int64_t)n_feat * n_lag;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n + p; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < n) {
      const int s = (int)(i / n_time), t = (int)(i % n_time);
      const int64_t tr = trial_idx ? trial_idx[s] : s;
      row_off[i] = tr * trial_stride + t;
      if (yrow_off) yrow_off[i] = tr * y_trial_stride + t;
    } else {
      const int64_t j = i - n;
      const int f = (int)(j / n_lag), w = (int)(j % n_lag);
      const int64_t ft = feat_idx ? feat_idx[f] : f;
      col_off[j] = ft * feat_stride + w;
    }
  }
}

__global__ void iota_offsets_kernel(int64_t* out, int64_t n, int64_t stride) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = i * stride;
}

// ------------------------------------------------------------------------------------------------
// CUDA-core contraction with the same separable-gather operands as the tensor-core core.
// fp64 path (reference tests run in fp64) and cross-check for the tcgen05 kernel.
// 64x64 tile, 16-deep K slabs, 256 threads, 4x4 outputs per thread; split-K partials over blockIdx.z.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct GOp {
  const T* base; const int64_t* roff; const int64_t* koff; int rows;
};

template <typename T>
__global__ void __launch_bounds__(256) simt_contract_kernel(GOp<T> A, GOp<T> B, int M, int N, int K, T* C, int64_t ldc,
                                                            int64_t split_stride, int symmetric) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ T As[TK][TM + 1];
  __shared__ T Bs[TK][TN + 1];
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  if (symmetric && n0 > m0 + TM - 1) return;
  const int nk = (K + TK - 1) / TK;
  const int per = (nk + gridDim.z - 1) / gridDim.z;
  const int kb0 = blockIdx.z * per, kb1 = min(nk, kb0 + per);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  typename Acc<T>::type acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0;
  for (int kb = kb0; kb < kb1; ++kb) {
    const int k0 = kb * TK;
    for (int e = threadIdx.x; e < TM * TK; e += 256) {
      const int r = e % TM, kk = e / TM;
      const int gr = m0 + r, gk = k0 + kk;
      As[kk][r] = (gr < A.rows && gk < K) ? A.base[A.roff[gr] + A.koff[gk]] : (T)0;
    }
    for (int e = threadIdx.x; e < TN * TK; e += 256) {
      const int r = e % TN, kk = e / TN;
      const int gr = n0 + r, gk = k0 + kk;
      Bs[kk][r] = (gr < B.rows && gk < K) ? B.base[B.roff[gr] + B.koff[gk]] : (T)0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += (typename Acc<T>::type)a[i] * (typename Acc<T>::type)b[j];
    }
    __syncthreads();
  }
  T* Cz = C + (int64_t)blockIdx.z * split_stride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      Cz[(int64_t)gm * ldc + gn] = (T)acc[i][j];
      if (symmetric && gn < M && gm < N) {
        const int mt0 = (gn / TM) * TM, nt0 = (gm / TN) * TN;
        if (nt0 > mt0 + TM - 1) Cz[(int64_t)gn * ldc + gm] = (T)acc[i][j];
      }
    }
  }
}

// out[m][n] = sum_z part[z][m][n]   (round-to-nearest adds of split-K partials, fixed order)
template <typename T>
__global__ void reduce_splits_kernel(const T* __restrict__ part, int splits, int64_t split_stride, int M, int N,
                                     int64_t ld_part, T* out, int64_t ld_out, int block_cols = 0, int64_t block_stride = 0) {
  const int64_t total = (int64_t)M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / N, n = i - m * N;
    typename Acc<T>::type s = 0;
    for (int z = 0; z < splits; ++z) s += part[z * split_stride + m * ld_part + n];
    if (block_cols > 0) out[(n / block_cols) * block_stride + m * block_cols + (n % block_cols)] = (T)s;   // column-blocked C
    else out[m * ld_out + n] = (T)s;
  }
}

// ------------------------------------------------------------------------------------------------
// column / target sums of the implicit design (for centring: ridgecv.py:81-88)
// one thread per column j, consecutive j -> consecutive lags -> coalesced; fp64 accumulation.
// grid.y splits the rows; partial sums are combined by sums_finish_kernel.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void col_sums_kernel(const T* __restrict__ base, const int64_t* __restrict__ row_off,
                                const int64_t* __restrict__ col_off, int64_t n, int64_t p, double* partial) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p) return;
  const int64_t per = (n + gridDim.y - 1) / gridDim.y;
  const int64_t i0 = blockIdx.y * per, i1 = min(n, i0 + per);
  const T* src = base + col_off[j];
  double s = 0;
  for (int64_t i = i0; i < i1; ++i) s += (double)src[row_off[i]];
  partial[(int64_t)blockIdx.y * p + j] = s;
}

template <typename T>
__global__ void sums_finish_kernel(const double* __restrict__ partial, int parts, int64_t p, double inv_n, int enabled,
                                   T* mean_out, double* mean64) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p) return;
  double s = 0;
  for (int z = 0; z < parts; ++z) s += partial[(int64_t)z * p + j];
  const double m = enabled ? s * inv_n : 0.0;
  mean_out[j] = (T)m;
  mean64[j] = m;
}

// G[i][j] -= n * mu_i * mu_j ;  XtY[j][f] -= n * mu_j * ybar_f    (centred moments from raw ones)
template <typename T>
__global__ void centre_moments_kernel(T* G, int64_t p, T* XtY, int F, const double* __restrict__ mu,
                                      const double* __restrict__ ybar, double n) {
  const int64_t totalG = p * p, total = totalG + p * F;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    if (e < totalG) {
      const int64_t i = e / p, j = e - i * p;
      G[e] = (T)((double)G[e] - n * mu[i] * mu[j]);
    } else {
      const int64_t q = e - totalG;
      const int64_t j = q / F, f = q - j * F;
      XtY[q] = (T)((double)XtY[q] - n * mu[j] * ybar[f]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// One-sided (Hestenes) Jacobi eigensolver on the rows of a symmetric matrix, batched.
//   B <- J^T B, Vt <- J^T Vt (Vt starts as I).  At convergence the rows of B are mutually orthogonal,
//   B = Lambda Vt, so row k of Vt is an eigenvector and lambda_k = <B_k, Vt_k>.
// One thread-block cluster per matrix; each round-robin step's p/2 disjoint row pairs are spread over
// the warps of the cluster; cluster barriers separate steps.  Rows live in global memory (L2).
// Replaces torch.linalg.svd / eigh of ridgecv.py:144,218 on the centred Gram.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024) jacobi_rows_kernel(T* Aall, T* Vall, T* lam_all, int p, int max_sweeps,
                                                           double tol, unsigned int* counters, int cluster_size) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int mat = blockIdx.x / cluster_size;
  T* B = Aall + (int64_t)mat * p * p;
  T* V = Vall + (int64_t)mat * p * p;
  T* lam = lam_all + (int64_t)mat * p;
  unsigned int* cnt = counters + (int64_t)mat * max_sweeps;
  const int warps_per_block = blockDim.x >> 5;
  const int gw = crank * warps_per_block + (threadIdx.x >> 5);
  const int nw = cluster_size * warps_per_block;
  const int lane = threadIdx.x & 31;
  const int m = p + (p & 1);

  for (int64_t e = (int64_t)crank * blockDim.x + threadIdx.x; e < (int64_t)p * p; e += (int64_t)cluster_size * blockDim.x) {
    const int64_t i = e / p, j = e - i * p;
    V[e] = (i == j) ? (T)1 : (T)0;
  }
  __threadfence();
  cluster.sync();

  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    unsigned int my_rot = 0;
    for (int step = 0; step < m - 1; ++step) {
      for (int q = gw; q < m / 2; q += nw) {
        int i, j;
        if (q == 0) { i = m - 1; j = step; }
        else { i = (step + q) % (m - 1); j = (step - q + (m - 1)) % (m - 1); }
        if (i >= p || j >= p) continue;
        T* bi = B + (int64_t)i * p; T* bj = B + (int64_t)j * p;
        double aa = 0, bb = 0, gg = 0;
        for (int k = lane; k < p; k += 32) {
          const double x = (double)__ldcg(bi + k), y = (double)__ldcg(bj + k);
          aa += x * x; bb += y * y; gg += x * y;
        }
        aa = warp_sum(aa); bb = warp_sum(bb); gg = warp_sum(gg);
        if (aa * bb > 0.0 && fabs(gg) > tol * sqrt(aa * bb)) {
          const double zeta = (bb - aa) / (2.0 * gg);
          const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
          T* vi = V + (int64_t)i * p; T* vj = V + (int64_t)j * p;
          for (int k = lane; k < p; k += 32) {
            const double x = (double)__ldcg(bi + k), y = (double)__ldcg(bj + k);
            __stcg(bi + k, (T)(c * x - s * y));
            __stcg(bj + k, (T)(s * x + c * y));
            const double u = (double)__ldcg(vi + k), w = (double)__ldcg(vj + k);
            __stcg(vi + k, (T)(c * u - s * w));
            __stcg(vj + k, (T)(s * u + c * w));
          }
          if (lane == 0) ++my_rot;
        }
      }
      __threadfence();
      cluster.sync();
    }
    if (lane == 0 && my_rot) atomicAdd(&cnt[sweep], my_rot);
    __threadfence();
    cluster.sync();
    const unsigned int total = __ldcg(&cnt[sweep]);
    if (total == 0) break;
  }
  for (int k = gw; k < p; k += nw) {
    const T* bk = B + (int64_t)k * p; const T* vk = V + (int64_t)k * p;
    double s = 0;
    for (int c = lane; c < p; c += 32) s += (double)__ldcg(bk + c) * (double)__ldcg(vk + c);
    s = warp_sum(s);
    if (lane == 0) lam[k] = (T)(s > 0 ? s : 0);
  }
}

// ------------------------------------------------------------------------------------------------
// LOO sweep pieces (ridgecv.py:230-251 restated in the primal basis; SURVEY.md section 8 a-note)
// ------------------------------------------------------------------------------------------------
// w[a][k] = 1 / (lam_k + alpha_a)
template <typename T>
__global__ void ridge_weights_kernel(const T* __restrict__ lam, const T* __restrict__ alphas, int A, int p, T* w) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (int64_t)A * p; e += (int64_t)gridDim.x * blockDim.x) {
    const int a = (int)(e / p), k = (int)(e - (int64_t)a * p);
    w[e] = (T)(1.0 / ((double)lam[k] + (double)alphas[a]));
  }
}

// theta_t[(a,f)][k] = w[a][k] * Vtb[k][f]
template <typename T>
__global__ void theta_kernel(const T* __restrict__ w, const T* __restrict__ Vtb, int A, int F, int p, T* theta_t) {
  const int64_t total = (int64_t)A * F * p;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t af = e / p;
    const int k = (int)(e - af * p);
    const int a = (int)(af / F), f = (int)(af - (int64_t)a * F);
    theta_t[e] = w[(int64_t)a * p + k] * Vtb[(int64_t)k * F + f];
  }
}

// out[r] = sum_j M[r][j] * v[j]   (row-major M, one warp per row; fp64 v)
template <typename T>
__global__ void matvec_rows_kernel(const T* __restrict__ M, int64_t rows, int64_t cols, const double* __restrict__ v,
                                   double* out) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  double s = 0;
  for (int64_t j = lane; j < cols; j += 32) s += (double)M[r * cols + j] * v[j];
  s = warp_sum(s);
  if (lane == 0) out[r] = s;
}

// d[i][a] = 1 - c0 - sum_k (Z[i][k] - cz[k])^2 * w[a][k]
// block: 64 rows x all alphas; K swept in slabs of 64 staged through shared memory.
template <typename T>
__global__ void __launch_bounds__(256) leverage_kernel(const T* __restrict__ Z, int64_t n, int p, int64_t ldz,
                                                       const double* __restrict__ cz, const T* __restrict__ w, int A,
                                                       double c0, T* d) {
  constexpr int RT = 64, KT = (sizeof(T) == 8) ? 32 : 64, AMAX = 32;
  __shared__ T zs[RT][KT + 1];
  __shared__ T ws[AMAX][KT];
  const int64_t r0 = (int64_t)blockIdx.x * RT;
  const int r = threadIdx.x & 63, ag = threadIdx.x >> 6;  // 4 alpha groups
  for (int a0 = 0; a0 < A; a0 += AMAX) {
    const int na = min(AMAX, A - a0);
    double acc[AMAX / 4];
#pragma unroll
    for (int u = 0; u < AMAX / 4; ++u) acc[u] = 0;
    for (int k0 = 0; k0 < p; k0 += KT) {
      for (int e = threadIdx.x; e < RT * KT; e += 256) {
        const int rr = e / KT, kk = e % KT;
        const int64_t gi = r0 + rr; const int gk = k0 + kk;
        T z = 0;
        if (gi < n && gk < p) z = (T)((double)Z[gi * ldz + gk] - cz[gk]);
        zs[rr][kk] = z * z;
      }
      for (int e = threadIdx.x; e < na * KT; e += 256) {
        const int aa = e / KT, kk = e % KT;
        const int gk = k0 + kk;
        ws[aa][kk] = (gk < p) ? w[(int64_t)(a0 + aa) * p + gk] : (T)0;
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < AMAX / 4; ++u) {
        const int aa = ag + 4 * u;
        if (aa < na) {
          double s = 0;
#pragma unroll 8
          for (int kk = 0; kk < KT; ++kk) s += (double)(zs[r][kk] * ws[aa][kk]);
          acc[u] += s;
        }
      }
      __syncthreads();
    }
    if (r0 + r < n) {
#pragma unroll
      for (int u = 0; u < AMAX / 4; ++u) {
        const int aa = ag + 4 * u;
        if (aa < na) d[(r0 + r) * A + a0 + aa] = (T)(1.0 - c0 - acc[u]);
      }
    }
  }
}

// loss partials: for column c=(a,f) and a slab of rows:
//   e = (y[i][f] - ybar_f - R[i][c] + mb[c]) / d[i][a];   partial[slab][c] = sum_i e^2
template <typename T>
__global__ void __launch_bounds__(256) loo_loss_kernel(const T* __restrict__ R, int64_t n, int A, int F, int64_t ldr,
                                                       const T* __restrict__ ybase, const int64_t* __restrict__ yrow_off,
                                                       int64_t y_feat_stride, const double* __restrict__ ybar,
                                                       const double* __restrict__ mb, const T* __restrict__ d,
                                                       double* partial) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int AF = A * F;
  const int64_t per = (n + gridDim.y - 1) / gridDim.y;
  const int64_t i0 = blockIdx.y * per, i1 = min(n, i0 + per);
  if (c >= AF) return;
  const int a = c / F, f = c - a * F;
  const double yb = ybar[f], m = mb[c];
  const T* ycol = ybase + (int64_t)f * y_feat_stride;
  double s = 0;
  for (int64_t i = i0; i < i1; ++i) {
    const double e = ((double)ycol[yrow_off[i]] - yb - (double)R[i * ldr + c] + m) / (double)d[i * A + a];
    s += e * e;
  }
  partial[(int64_t)blockIdx.y * AF + c] = s;
}

// loss[a][f] = sum_slabs / n ; then first-minimum selection (ridgecv.py:253-256, 271-274);
// single block.  per_target: best[f] = argmin_a loss[a][f]; else best[0] = argmin_a mean_f loss[a][f]
template <typename T>
__global__ void select_alpha_kernel(const double* __restrict__ partial, int slabs, int A, int F, double inv_n,
                                    const T* __restrict__ alphas, int per_target, T* loss_out, int32_t* best,
                                    T* alpha_out, double* loss64) {
  const int AF = A * F;
  for (int c = threadIdx.x; c < AF; c += blockDim.x) {
    double s = 0;
    for (int z = 0; z < slabs; ++z) s += partial[(int64_t)z * AF + c];
    s *= inv_n;
    loss64[c] = s;
    if (loss_out) loss_out[c] = (T)s;
  }
  __syncthreads();
  if (per_target) {
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
      // first minimum; a non-finite loss (numerically singular design at a vanishing alpha) never wins over a finite one
      int b = 0; T bv = (T)loss64[f];
      for (int a = 1; a < A; ++a) {
        const T v = (T)loss64[a * F + f];
        if (v < bv || (!(bv == bv) && v == v)) { bv = v; b = a; }
      }
      best[f] = b;
      alpha_out[f] = alphas[b];
    }
  } else if (threadIdx.x == 0) {
    int b = 0; T bv = 0;
    for (int a = 0; a < A; ++a) {
      // mean over targets in working precision T, like losses.mean(1) on a T tensor
      double s = 0;
      for (int f = 0; f < F; ++f) s += (double)(T)loss64[a * F + f];
      const T v = (T)(s / F);
      if (a == 0 || v < bv || (!(bv == bv) && v == v)) { bv = v; b = a; }
    }
    best[0] = b;
    alpha_out[0] = alphas[b];
  }
}

// coef[f][j] = beta_all[(best_f, f)][j];  intercept[f] = ybar_f - mb[(best_f, f)]
template <typename T>
__global__ void gather_coef_kernel(const T* __restrict__ beta_all, int F, int64_t p, const int32_t* __restrict__ best,
                                   int per_target, const double* __restrict__ ybar, const double* __restrict__ mb,
                                   int fit_intercept, T* coef, T* intercept) {
  const int64_t total = (int64_t)F * p;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(e / p);
    const int64_t j = e - (int64_t)f * p;
    const int a = per_target ? best[f] : best[0];
    coef[e] = beta_all[((int64_t)a * F + f) * p + j];
    if (j == 0) intercept[f] = fit_intercept ? (T)(ybar[f] - mb[(int64_t)a * F + f]) : (T)0;
  }
}

// yhat[(s*F + f)*T + t] = P[(s*T + t)][f] + intercept[f]    (timedelayed.py:531-533 reshape/swapaxes)
template <typename T>
__global__ void predict_layout_kernel(const T* __restrict__ P, int64_t n_trials, int F, int Tn, int64_t ldp,
                                      const T* __restrict__ intercept, T* yhat) {
  __shared__ T tile[32][33];
  const int s = blockIdx.z;
  const int t0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int t = t0 + j, f = f0 + threadIdx.x;
    if (t < Tn && f < F) tile[j][threadIdx.x] = P[((int64_t)s * Tn + t) * ldp + f];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int f = f0 + j, t = t0 + threadIdx.x;
    if (t < Tn && f < F) yhat[((int64_t)s * F + f) * Tn + t] = tile[threadIdx.x][j] + (intercept ? intercept[f] : (T)0);
  }
}

// ------------------------------------------------------------------------------------------------
// Fused scoring: y, yhat viewed as [R][K]; each thread owns one kept cell k and walks the R reduced
// samples (coalesced across k).  One read of y / yhat per pass serves both metrics.
//   r2      = 1 - sum (y - yhat)^2 / sum (y - y_b)^2, 0 where the denominator <= 0   (math/r2.py:92-100)
//   pearson = sum dx dy / sqrt(sum dx^2 sum dy^2) with dx, dy centred                  (math/pearsonr.py:55-56)
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// average ranks along the reduced axis (math/rank.py:62-112): x viewed as [R][K], out[r][k] = 1 + #{x_j < x_r}
// + (#{x_j == x_r} - 1) / 2 over column k.  One thread per cell, O(R^2) comparisons from a register cache when
// R <= 64 (the trials-first reduce of the metrics), strided re-reads otherwise.
// ------------------------------------------------------------------------------------------------
template <typename T, int RC>
__global__ void rank_kernel(const T* __restrict__ x, int64_t R, int64_t K, T* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  if (RC > 0) {
    T v[RC > 0 ? RC : 1];
#pragma unroll
    for (int r = 0; r < RC; ++r) v[r] = (r < R) ? __ldg(x + r * K + k) : (T)0;
#pragma unroll
    for (int r = 0; r < RC; ++r) {
      if (r < R) {
        int less = 0, equal = 0;
#pragma unroll
        for (int j = 0; j < RC; ++j) {
          if (j < R) { less += v[j] < v[r]; equal += v[j] == v[r]; }
        }
        out[r * K + k] = (T)(1.0 + less + 0.5 * (equal - 1));
      }
    }
  } else {
    for (int64_t r = 0; r < R; ++r) {
      const T xr = x[r * K + k];
      int64_t less = 0, equal = 0;
      for (int64_t j = 0; j < R; ++j) { const T xj = x[j * K + k]; less += xj < xr; equal += xj == xr; }
      out[r * K + k] = (T)(1.0 + (double)less + 0.5 * (double)(equal - 1));
    }
  }
}

// One thread per kept cell k; the reduced samples r are the slow axis, so every load is coalesced across the warp.
// RC > 0: R <= RC samples are held in registers (all loads issued up front, each value read from memory once);
// RC == 0: generic two-pass loop for long reduce dims.  fp64 accumulators, results rounded to T like the reference's
// fp32 tensor ops.
template <typename T, int RC>
__global__ void score_kernel(const T* __restrict__ y, const T* __restrict__ yh, const T* __restrict__ yb, int64_t R,
                             int64_t K, T* r2, T* pr) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double sy = 0, sh = 0, num = 0, den = 0, sxy = 0, sxx = 0, syy = 0;
  if (RC > 0) {
    T a[RC > 0 ? RC : 1], b[RC > 0 ? RC : 1];
#pragma unroll
    for (int r = 0; r < RC; ++r) {
      a[r] = (r < R) ? __ldg(y + r * K + k) : (T)0;
      b[r] = (r < R) ? __ldg(yh + r * K + k) : (T)0;
    }
    double sy2 = 0, sh2 = 0;
#pragma unroll
    for (int r = 0; r < RC; r += 2) { sy += (double)a[r]; sh += (double)b[r]; if (r + 1 < RC) { sy2 += (double)a[r + 1]; sh2 += (double)b[r + 1]; } }
    sy += sy2; sh += sh2;
    const T my = (T)(sy / R), mh = (T)(sh / R);
    const T base = yb ? yb[k] : my;
#pragma unroll
    for (int r = 0; r < RC; ++r) {
      if (r < R) {
        const T e = a[r] - b[r], g = a[r] - base, dx = a[r] - my, dy = b[r] - mh;
        num += (double)(e * e); den += (double)(g * g);
        sxy += (double)(dx * dy); sxx += (double)(dx * dx); syy += (double)(dy * dy);
      }
    }
  } else {
    for (int64_t r = 0; r < R; ++r) { sy += (double)y[r * K + k]; sh += (double)yh[r * K + k]; }
    const T my = (T)(sy / R), mh = (T)(sh / R);
    const T base = yb ? yb[k] : my;
    for (int64_t r = 0; r < R; ++r) {
      const T av = y[r * K + k], bv = yh[r * K + k];
      const T e = av - bv, g = av - base, dx = av - my, dy = bv - mh;
      num += (double)(e * e); den += (double)(g * g);
      sxy += (double)(dx * dy); sxx += (double)(dx * dx); syy += (double)(dy * dy);
    }
  }
  if (r2) r2[k] = ((T)den > (T)0) ? (T)1 - (T)num / (T)den : (T)0;
  if (pr) pr[k] = (T)sxy / (T)sqrt((double)((T)sxx * (T)syy));
}


// ---- fused r2 + pearson, vectorised and warp-shuffle-reduced (math/r2.py:57-101, math/pearsonr.py:35-56) -----------------
// y, yh: [R][K] with the reduced samples as the slow axis.  A lane owns one 16-byte group of cells and walks the rows; RS
// lanes-groups of a warp take interleaved rows (RS = 1: a warp reads 512 contiguous bytes per row; RS = 4: 128 bytes of four
// rows, for small problems that would otherwise leave the SMs empty) and are combined with xor-shuffles at the end.  Rows are
// unrolled by four: eight independent 16-byte loads in flight per lane, from at most 8 RS different rows -- measured with ncu
// (profiles/r02_notes.md): what bounds this kernel is memory latency, and that latency explodes once a warp has loads to more
// than a few dozen different 2 MB pages in flight (the rows of one cell group are K * 4 bytes apart), so the row unroll is
// kept short and the occupancy high instead.  Moments are taken about a pivot (the first sample of the cell: a constant
// series then has exactly zero variance and scores NaN like the reference's 0 / 0), summed in T over blocks of 32 rows and
// folded into a second set of accumulators; the centred sums follow in fp64 as S_xy - S_x S_y / R.  One pass, each input byte
// read from HBM once.
template <typename T> struct ScoreVec;
template <> struct ScoreVec<float> { using V = float4; static constexpr int W = 4; };
template <> struct ScoreVec<double> { using V = double2; static constexpr int W = 2; };
__device__ __forceinline__ void vec_get(const float4& v, float (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ void vec_get(const double2& v, double (&o)[2]) { o[0] = v.x; o[1] = v.y; }

template <typename T, int RS, int U>
__global__ void __launch_bounds__(256, U == 2 ? 3 : 2) score_vec_kernel(const T* __restrict__ y, const T* __restrict__ yh, const T* __restrict__ yb,
                                                        int64_t R, int64_t K, T* __restrict__ r2, T* __restrict__ pr) {
  using V = typename ScoreVec<T>::V;
  constexpr int W = ScoreVec<T>::W;
  constexpr int CGW = 32 / RS;                                 // column groups per warp
  const int lane = threadIdx.x & 31, cg = lane % CGW, rs = lane / CGW;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t k0 = (warp * CGW + cg) * W;                    // first cell of this lane (K % W == 0)
  const bool ok = k0 < K;
  T base[W], py0[W], ph0[W];
  T ty[W], th[W], tyy[W], thh[W], tyh[W], tee[W], tgg[W];      // current block of rows
  T sy[W], sh[W], syy[W], shh[W], syh[W], see[W], sgg[W];      // folded blocks
#pragma unroll
  for (int u = 0; u < W; ++u) {
    base[u] = py0[u] = ph0[u] = (T)0;
    ty[u] = th[u] = tyy[u] = thh[u] = tyh[u] = tee[u] = tgg[u] = (T)0;
    sy[u] = sh[u] = syy[u] = shh[u] = syh[u] = see[u] = sgg[u] = (T)0;
  }
  if (ok) {
    if (yb) vec_get(__ldg(reinterpret_cast<const V*>(yb + k0)), base);
    vec_get(__ldg(reinterpret_cast<const V*>(y + k0)), py0);            // pivots: the cell's first sample
    vec_get(__ldg(reinterpret_cast<const V*>(yh + k0)), ph0);
  }
  auto add = [&](const V& av, const V& bv) {
    T a[W], b[W];
    vec_get(av, a); vec_get(bv, b);
#pragma unroll
    for (int u = 0; u < W; ++u) {
      const T e = a[u] - b[u], g = a[u] - base[u], da = a[u] - py0[u], db = b[u] - ph0[u];
      ty[u] += da; th[u] += db;
      tyy[u] = fma(da, da, tyy[u]); thh[u] = fma(db, db, thh[u]); tyh[u] = fma(da, db, tyh[u]);
      tee[u] = fma(e, e, tee[u]); tgg[u] = fma(g, g, tgg[u]);
    }
  };
  auto fold = [&]() {
#pragma unroll
    for (int u = 0; u < W; ++u) {
      sy[u] += ty[u]; sh[u] += th[u]; syy[u] += tyy[u]; shh[u] += thh[u]; syh[u] += tyh[u]; see[u] += tee[u]; sgg[u] += tgg[u];
      ty[u] = th[u] = tyy[u] = thh[u] = tyh[u] = tee[u] = tgg[u] = (T)0;
    }
  };
  if (ok) {
    const V* py = reinterpret_cast<const V*>(y + k0);
    const V* ph = reinterpret_cast<const V*>(yh + k0);
    const int64_t stride = K / W;                              // row pitch in vectors
    int64_t r = rs;
    int blk = 0;
    for (; r + (U - 1) * RS < R; r += U * RS) {
      V a[U], b[U];
#pragma unroll
      for (int i = 0; i < U; ++i) { a[i] = __ldg(py + (r + i * RS) * stride); b[i] = __ldg(ph + (r + i * RS) * stride); }
#pragma unroll
      for (int i = 0; i < U; ++i) add(a[i], b[i]);
      if (++blk == 32 / U) { fold(); blk = 0; }
    }
    for (; r < R; r += RS) add(__ldg(py + r * stride), __ldg(ph + r * stride));
    fold();
  }
#pragma unroll
  for (int o = CGW; o < 32; o <<= 1)
#pragma unroll
    for (int u = 0; u < W; ++u) {
      sy[u] += __shfl_xor_sync(0xffffffffu, sy[u], o); sh[u] += __shfl_xor_sync(0xffffffffu, sh[u], o);
      syy[u] += __shfl_xor_sync(0xffffffffu, syy[u], o); shh[u] += __shfl_xor_sync(0xffffffffu, shh[u], o);
      syh[u] += __shfl_xor_sync(0xffffffffu, syh[u], o); see[u] += __shfl_xor_sync(0xffffffffu, see[u], o);
      sgg[u] += __shfl_xor_sync(0xffffffffu, sgg[u], o);
    }
  if (ok && rs == 0) {
    const double inv = 1.0 / (double)R;
#pragma unroll
    for (int u = 0; u < W; ++u) {
      const double dsy = sy[u], dsh = sh[u];
      const double cxx = (double)syy[u] - dsy * dsy * inv, chh = (double)shh[u] - dsh * dsh * inv, cxy = (double)syh[u] - dsy * dsh * inv;
      const double den = yb ? (double)sgg[u] : cxx;            // without a baseline: the mean of y (r2.py:80-81)
      if (r2) r2[k0 + u] = ((T)den > (T)0) ? (T)1 - (T)see[u] / (T)den : (T)0;
      // a constant series has all pivot moments exactly zero: 0 / 0 = NaN as in the reference (pearsonr.py:55-56, no guard)
      if (pr) pr[k0 + u] = (T)cxy / (T)sqrt((double)((T)fmax(cxx, 0.0) * (T)fmax(chh, 0.0)));
    }
  }
}

// ---- Wu-Lin pairwise coupling of one-versus-one probabilities (estimators/classifier.py:527-653) ----------------------
// One thread per sample, K <= COUPLING_MAX_K classes, double arithmetic.  The reference builds Q and r with in-place adds
// over index tensors that repeat (Q[:, j_i, j_i] += ..., r[:, j_i] += ... with j_i = [1, 2, 2, 3, 3, 3, ...]): such an update
// keeps the LAST write per index instead of accumulating, so the diagonal of Q and r receive, for class j, only the
// contributions of the pairs (j, j - 1) [as the larger index] and (K - 1, j) [as the smaller]; the off-diagonal entries are
// per pair.  That is what the reference computes, and what is reproduced here.
// The reference stops all samples at the first iteration whose largest change over the batch is < tol: iteration `it` is
// a launch that returns at once when an earlier iteration's maximum (maxdiff[j], j < it) was below tol.
constexpr int COUPLING_MAX_K = 16;
template <typename T>
__global__ void coupling_step_kernel(const T* __restrict__ P, int N, int K, double eps, double tol, int it,
                                     const double* __restrict__ p_in, double* __restrict__ p_out, double* maxdiff) {
  for (int j = 0; j < it; ++j)
    if (maxdiff[j] < tol) return;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  constexpr int KM = COUPLING_MAX_K;
  double A[KM + 1][KM + 2], p[KM];
  const T* Pn = P + (int64_t)n * K * K;
  for (int j = 0; j < K; ++j) p[j] = it == 0 ? 1.0 / K : p_in[(int64_t)n * K + j];
  for (int a = 0; a <= K; ++a)
    for (int b = 0; b <= K + 1; ++b) A[a][b] = 0.0;
  auto terms = [&](int j, int k, double& pij, double& pji, double& w, double& denom) {
    pij = (double)Pn[j * K + k]; pji = (double)Pn[k * K + j];
    const double pr = pij * pji;
    w = 1.0 / (pr < eps ? eps : pr);
    denom = p[j] + p[k] + eps;
  };
  double pij, pji, w, denom;
  for (int j = 1; j < K; ++j) {                       // "relative to j": the last pair with the larger index j is (j, j - 1)
    terms(j, j - 1, pij, pji, w, denom);
    A[j][j] = w * pji / (denom * denom);
    A[j][K + 1] = w * (pij - p[j] / denom);
  }
  for (int k = 0; k < K - 1; ++k) {                   // "relative to k": the last pair with the smaller index k is (K - 1, k)
    terms(K - 1, k, pij, pji, w, denom);
    A[k][k] += w * pij / (denom * denom);
    A[k][K + 1] += w * (pji - p[k] / denom);
  }
  for (int j = 1; j < K; ++j)
    for (int k = 0; k < j; ++k) {
      terms(j, k, pij, pji, w, denom);
      A[j][k] -= w * pji / (denom * denom);
      A[k][j] -= w * pij / (denom * denom);
    }
  for (int j = 0; j < K; ++j) { A[j][j] += 1e-9; A[j][K] = 1.0; A[K][j] = 1.0; }      // KKT system [[Q + 1e-9 I, 1], [1^T, 0]]
  A[K][K] = 0.0; A[K][K + 1] = 0.0;
  const int M = K + 1;
  for (int c = 0; c < M; ++c) {                       // Gaussian elimination with partial pivoting (torch.linalg.solve)
    int piv = c;
    double best = fabs(A[c][c]);
    for (int rr = c + 1; rr < M; ++rr) if (fabs(A[rr][c]) > best) { best = fabs(A[rr][c]); piv = rr; }
    if (piv != c) for (int b = c; b <= M; ++b) { const double t = A[c][b]; A[c][b] = A[piv][b]; A[piv][b] = t; }
    const double inv = 1.0 / A[c][c];
    for (int rr = c + 1; rr < M; ++rr) {
      const double f = A[rr][c] * inv;
      if (f != 0.0) for (int b = c; b <= M; ++b) A[rr][b] -= f * A[c][b];
    }
  }
  double x[KM + 1];
  for (int c = M - 1; c >= 0; --c) {
    double s = A[c][M];
    for (int b = c + 1; b < M; ++b) s -= A[c][b] * x[b];
    x[c] = s / A[c][c];
  }
  double pn[KM], sum = 0.0, diff = 0.0;
  for (int j = 0; j < K; ++j) { pn[j] = p[j] + x[j]; if (pn[j] < 0.0) pn[j] = 0.0; sum += pn[j]; }
  for (int j = 0; j < K; ++j) {
    pn[j] = sum > eps ? pn[j] / sum : 1.0 / K;
    diff += fabs(pn[j] - p[j]);
    p_out[(int64_t)n * K + j] = pn[j];
  }
  atomicMax(reinterpret_cast<unsigned long long*>(maxdiff + it), (unsigned long long)__double_as_longlong(diff));   // diff >= 0: bit order = value order
}
// the estimate of the iteration that stopped the loop (or of the last one)
template <typename T>
__global__ void coupling_finish_kernel(const double* __restrict__ buf0, const double* __restrict__ buf1, const double* __restrict__ maxdiff,
                                       int max_iter, double tol, int64_t total, T* __restrict__ out) {
  int c = max_iter - 1;
  for (int j = 0; j < max_iter; ++j) if (maxdiff[j] < tol) { c = j; break; }
  const double* src = ((c + 1) & 1) ? buf1 : buf0;          // iteration it reads buf[it & 1], writes buf[(it + 1) & 1]
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) out[e] = (T)src[e];
}


// ---- statistics of a predictor subset as a slice of the full-predictor statistics --------------------------------------
// Columns of the lag-expanded design are (feature, lag): the subset `feat_idx` of the features selects
// cols[c * W + w] = feat_idx[c] * W + w.  G_sub = G[cols][:, cols], XtY_sub = XtY[cols], mu_sub = mu[cols]
// (hierarchical.py / shapley.py refit every predictor set from scratch; SURVEY 8 a-note: the subset's Gram is a principal
// sub-block of the fold's full Gram).  One pass: each output element is read once and written once.
template <typename T>
__global__ void gather_stats_kernel(const T* __restrict__ G, const T* __restrict__ XtY, const T* __restrict__ mu, int64_t p_full, int F,
                                    const int32_t* __restrict__ feat_idx, int n_feat, int W, T* __restrict__ Gs, T* __restrict__ XtYs,
                                    T* __restrict__ mus) {
  const int64_t ps = (int64_t)n_feat * W;
  const int64_t total = ps * ps + ps * F + ps;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    if (e < ps * ps) {
      const int64_t r = e / ps, c = e - r * ps;
      const int64_t rf = (int64_t)feat_idx[r / W] * W + r % W, cf = (int64_t)feat_idx[c / W] * W + c % W;
      Gs[e] = G[rf * p_full + cf];
    } else if (e < ps * ps + ps * F) {
      const int64_t q = e - ps * ps, r = q / F, f = q - r * F;
      XtYs[q] = XtY[((int64_t)feat_idx[r / W] * W + r % W) * F + f];
    } else {
      const int64_t r = e - ps * ps - ps * F;
      mus[r] = mu[(int64_t)feat_idx[r / W] * W + r % W];
    }
  }
}


// out[i] = tab[i * stride]
__global__ void stride_pick_kernel(const int64_t* __restrict__ tab, int stride, int n, int64_t* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = tab[(int64_t)i * stride];
}

}  // namespace mvb
