// gram_recurrence.cuh -- X^T X of a lag-expanded design from its displacement structure: O(n p C + p^2 N) instead of O(n p^2).
//
//   G[(c,w),(c',w')] = sum_{trial} sum_{t=0}^{T-1} x_c[t+w] x_c'[t+w']          (x = the padded stimulus of one trial)
//
// Moving one step along a diagonal of the (c, c') block only exchanges the first term of the sum over t for one new last term:
//
//   G[(c,w+1),(c',w'+1)] = G[(c,w),(c',w')] + sum_trial ( x_c[w+T] x_c'[w'+T] - x_c[w] x_c'[w'] )
//
// so the W x W block follows from its first row and first column -- G[(c,0),(c',w')], which is one skinny contraction
// B0 = X_0^T X (the C lag-0 columns against all p columns, 1 % of the Gram's flops, on the tensor-core kernel), first column by
// symmetry -- and from Delta = A_1^T B_1 - A_0^T B_0, a K = 2 n_trials outer-product sum of the heads x[0..W) and tails x[T..T+W) of the
// two series.  The reference materialises the design and multiplies (timedelayed.py:394-431, ridgecv.py); MVPy's own
// ReceptiveField gets the same matrices from FFT correlations plus edge corrections (receptivefield.py:1046-1116) -- this is
// that idea without the FFT.  One CTA per (c, c') block: Delta in fp32 registers (7 x 7 per thread, trials streamed through shared
// memory in chunks of 16), then one thread per diagonal runs the prefix sum in fp64 and the block is written row by row.
// Exact up to rounding: the fp64 prefix adds 201 fp32-rounded corrections of relative size N/n to an entry.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mvb {
namespace gr {

constexpr int MAXW = 224;                 // lags per feature the shared-memory block can hold (7 x 32)
constexpr int TCH = 16;                   // trials per shared-memory chunk
constexpr int NTHR = 1024;

inline size_t smem_bytes(int W) { return ((size_t)W * (W + 1) + 4 * TCH * MAXW) * sizeof(float); }
inline bool usable(int n_lag) { return n_lag >= 2 && n_lag <= MAXW && smem_bytes(n_lag) <= 227 * 1024; }

// B0 [n_feat][p]: B0[c][c' W + w'] = G[(c,0),(c',w')] (uncentred).  G [p][ldg].
__global__ void __launch_bounds__(NTHR, 1) gram_blocks_kernel(const float* __restrict__ xp, const int32_t* __restrict__ trial_idx, int n_trials,
                                                              int64_t trial_stride, const int32_t* __restrict__ feat_idx, int n_feat,
                                                              int64_t feat_stride, int W, int Tn, const float* __restrict__ B0,
                                                              float* __restrict__ G, int64_t ldg) {
  extern __shared__ float sm[];
  const int c = blockIdx.y, cp = blockIdx.x;
  const int LD = W + 1;
  float* D = sm;                               // [W][W + 1]: Delta, then the finished block
  float* a0 = sm + (size_t)W * LD;             // [TCH][MAXW] heads of series c
  float* a1 = a0 + TCH * MAXW;                 // tails of series c (shifted by T)
  float* b0 = a1 + TCH * MAXW;                 // heads / tails of series c'
  float* b1 = b0 + TCH * MAXW;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* xa = xp + (int64_t)(feat_idx ? feat_idx[c] : c) * feat_stride;
  const float* xb = xp + (int64_t)(feat_idx ? feat_idx[cp] : cp) * feat_stride;
  float acc[7][7];
#pragma unroll
  for (int i = 0; i < 7; ++i)
#pragma unroll
    for (int j = 0; j < 7; ++j) acc[i][j] = 0.f;
  for (int t0 = 0; t0 < n_trials; t0 += TCH) {
    const int nt = min(TCH, n_trials - t0);
    __syncthreads();
    for (int e = threadIdx.x; e < TCH * MAXW; e += NTHR) {
      const int tr = e / MAXW, w = e - tr * MAXW;
      float va0 = 0.f, va1 = 0.f, vb0 = 0.f, vb1 = 0.f;
      if (tr < nt && w < W - 1) {                // Delta is needed for w, w' <= W - 2: x[w + T] stays inside the padded series
        const int64_t off = (int64_t)(trial_idx ? trial_idx[t0 + tr] : t0 + tr) * trial_stride + w;
        va0 = xa[off]; va1 = xa[off + Tn]; vb0 = xb[off]; vb1 = xb[off + Tn];
      }
      a0[e] = va0; a1[e] = va1; b0[e] = vb0; b1[e] = vb1;
    }
    __syncthreads();
    for (int tr = 0; tr < nt; ++tr) {
      float h0[7], h1[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) { h0[i] = a0[tr * MAXW + ty + 32 * i]; h1[i] = a1[tr * MAXW + ty + 32 * i]; }      // warp-uniform: broadcast
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const float v0 = b0[tr * MAXW + tx + 32 * j], v1 = b1[tr * MAXW + tx + 32 * j];
#pragma unroll
        for (int i = 0; i < 7; ++i) acc[i][j] = fmaf(h1[i], v1, fmaf(-h0[i], v0, acc[i][j]));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 7; ++i)
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int w = ty + 32 * i, wp = tx + 32 * j;
      if (w < W && wp < W) D[w * LD + wp] = acc[i][j];
    }
  __syncthreads();
  // one thread per diagonal: start on the block's first row (d < W: (0, d)) or first column ((d - W + 1, 0)), fp64 running sum
  for (int d = threadIdx.x; d < 2 * W - 1; d += NTHR) {
    int w = d < W ? 0 : d - W + 1, wp = d < W ? d : 0;
    double g = d < W ? (double)B0[(int64_t)c * ((int64_t)n_feat * W) + (int64_t)cp * W + wp]
                     : (double)B0[(int64_t)cp * ((int64_t)n_feat * W) + (int64_t)c * W + w];        // G[(c,w),(c',0)] = G[(c',0),(c,w)]
    for (; w < W && wp < W; ++w, ++wp) {
      const float delta = D[w * LD + wp];
      D[w * LD + wp] = (float)g;
      g += (double)delta;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < W * W; e += NTHR) {
    const int w = e / W, wp = e - w * W;
    G[((int64_t)c * W + w) * ldg + (int64_t)cp * W + wp] = D[w * LD + wp];
  }
}

inline cudaError_t launch(const float* xp, const int32_t* trial_idx, int n_trials, int64_t trial_stride, const int32_t* feat_idx, int n_feat,
                          int64_t feat_stride, int W, int Tn, const float* B0, float* G, int64_t ldg, cudaStream_t st) {
  const size_t smem = smem_bytes(W);
  cudaError_t e = cudaFuncSetAttribute(gram_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid(n_feat, n_feat);
  gram_blocks_kernel<<<grid, NTHR, smem, st>>>(xp, trial_idx, n_trials, trial_stride, feat_idx, n_feat, feat_stride, W, Tn, B0, G, ldg);
  return cudaGetLastError();
}

}  // namespace gr
}  // namespace mvb
