// jacobi_block.cuh -- block one-sided Jacobi eigensolver on the tensor-core contraction core (fp32).
//
// Replaces torch.linalg.svd / eigh of mvpy/estimators/ridgecv.py:144,218 for the centred Gram G (p x p).
//
// State: W = [B | Vt] (P x 2P, P = p padded to an even number of b-row blocks), B starts as G, Vt as I.
// One sweep visits every pair of row blocks once (round-robin tournament: all pairs of a step are
// disjoint, so a step is three batched launches):
//   1. S_q  = R_q R_q^T          Gram of the pair's 2b rows of B            (tcgen05 split-TF32, batched)
//   2. J_q  = eigenvectors(S_q)  one CTA per pair, one-sided Jacobi in shared memory, sorted
//   3. W'[rows_q] = J_q W[rows_q]  rotate the rows of B and Vt together      (tcgen05 split-TF32, batched)
// At convergence the rows of B are mutually orthogonal: B = Lambda Vt, row k of Vt is an eigenvector,
// lambda_k = <B_k, Vt_k>.  Convergence (no rotation needed anywhere in a sweep) is detected on the
// device; later launches see the flag and return at once, so the host never synchronises.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "tc_gemm.cuh"

namespace mvb {
namespace bj {

constexpr int BLK = 64;          // rows per block
constexpr int PR = 2 * BLK;      // rows per pair (= one MMA M tile)
constexpr int kMaxSweeps = 16;

struct Ctl {                     // device-side control block
  unsigned int done;             // 1 once a sweep needed no rotation
  unsigned int sweeps;           // sweeps executed when done was set (or max)
  unsigned int rot[kMaxSweeps];  // rotations counted per sweep
};

// W0 = [G padded | I]
__global__ void init_kernel(const float* __restrict__ G, int p, int P, float* W) {
  const int64_t total = (int64_t)P * 2 * P;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / (2 * P), c = e - r * 2 * P;
    float v = 0.f;
    if (c < P) { if (r < p && c < p) v = G[r * p + c]; }
    else if (c - P == r) v = 1.f;
    W[e] = v;
  }
}

// round-robin pairing tables for every step: rows[s][q][r] = element offset of the r-th row of pair q
__global__ void pair_tables_kernel(int nblk, int64_t ldw, int64_t* rows) {
  const int steps = nblk - 1, npairs = nblk / 2;
  const int64_t total = (int64_t)steps * npairs * PR;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e % PR);
    const int q = (int)((e / PR) % npairs);
    const int s = (int)(e / ((int64_t)PR * npairs));
    int bi, bjk;
    if (q == 0) { bi = nblk - 1; bjk = s; }
    else { bi = (s + q) % (nblk - 1); bjk = (s - q + (nblk - 1)) % (nblk - 1); }
    if (bi > bjk) { const int t = bi; bi = bjk; bjk = t; }
    const int blk = r < BLK ? bi : bjk;
    rows[e] = ((int64_t)blk * BLK + (r % BLK)) * ldw;
  }
}

// J tables: jrows[q][r] = q*PR*PR + r*PR ; iota
__global__ void misc_tables_kernel(int npairs, int64_t n_iota, int64_t* jrows, int64_t* iota) {
  const int64_t total = (int64_t)npairs * PR + n_iota;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    if (e < (int64_t)npairs * PR) jrows[e] = e * PR;
    else iota[e - (int64_t)npairs * PR] = e - (int64_t)npairs * PR;
  }
}

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One CTA per pair: eigenvectors of the pair's symmetric PSD 128 x 128 Gram S = R R^T, i.e. the rotation that
// makes the 128 rows R mutually orthogonal (implicit Hestenes).  Two-sided Jacobi in shared memory with the
// scale-invariant criterion |S_ij| <= tol sqrt(S_ii S_jj) (the cosine between rows i and j), 64 disjoint
// pairs per step, two per warp.  J is written with rows sorted by descending eigenvalue, which is what makes
// cyclic block Jacobi converge quickly; the orthogonality J loses to fp32 rounding is restored by the
// Newton-Schulz refresh of Vt in solve().  Pairs whose cosine exceeds tol_count in the first pass are counted as "work left" for this sweep.
constexpr int SLD = PR + 1;   // padded leading dimension of S (conflict-free column walks)
constexpr int SMALL_EIG_SMEM = PR * SLD * 4 + PR * PR * 4;

__global__ void __launch_bounds__(1024) small_eig_kernel(const float* __restrict__ S_all, float* __restrict__ J_all,
                                                         float tol_rot, float tol_count, int max_inner, Ctl* ctl, int sweep,
                                                         float* __restrict__ lam_out = nullptr,
                                                         const unsigned int* skip = nullptr) {
  if (ctl->done || (skip && *skip)) return;
  extern __shared__ __align__(16) unsigned char sm_raw[];
  float* J = reinterpret_cast<float*>(sm_raw);                         // [PR][PR]
  float* S = reinterpret_cast<float*>(sm_raw + PR * PR * 4);           // [PR][SLD]
  __shared__ float lam[PR];
  __shared__ int rank[PR];
  __shared__ unsigned int n_cnt;
  const float* Sg = S_all + (int64_t)blockIdx.x * PR * PR;
  float* Jg = J_all + (int64_t)blockIdx.x * PR * PR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < PR * PR; e += blockDim.x) {
    const int r = e / PR, c = e % PR;
    S[r * SLD + c] = 0.5f * (Sg[e] + Sg[c * PR + r]);   // symmetrise the tensor-core Gram
    J[e] = (r == c) ? 1.f : 0.f;
  }
  if (threadIdx.x == 0) n_cnt = 0;
  __syncthreads();
  // each warp owns pairs q = warp and warp + 32 of every step, in the row phase and in the column phase
  for (int inner = 0; inner < max_inner; ++inner) {
    int any_rot = 0;
    unsigned int cnt = 0;
    for (int step = 0; step < PR - 1; ++step) {
      int pi[2], pj[2];
      float pc[2], ps[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int q = warp + 32 * h;
        int i, j;
        if (q == 0) { i = PR - 1; j = step; }
        else { i = (step + q) % (PR - 1); j = (step - q + (PR - 1)) % (PR - 1); }
        pi[h] = i; pj[h] = j;
        const float aii = S[i * SLD + i], ajj = S[j * SLD + j], aij = S[i * SLD + j];   // broadcast reads
        float c = 1.f, sn = 0.f;
        const float scale = sqrtf(fabsf(aii * ajj));
        if (aii > 0.f && ajj > 0.f && fabsf(aij) > tol_rot * scale) {
          const float zeta = (ajj - aii) / (2.f * aij);
          const float t = (zeta >= 0.f ? 1.f : -1.f) / (fabsf(zeta) + sqrtf(1.f + zeta * zeta));
          c = 1.f / sqrtf(1.f + t * t);
          sn = c * t;
          any_rot = 1;
          if (inner == 0 && fabsf(aij) > tol_count * scale) ++cnt;
        }
        pc[h] = c; ps[h] = sn;
      }
      __syncwarp();
      // rows: (row_i, row_j) <- (c row_i - s row_j, s row_i + c row_j), for S and for J
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (ps[h] != 0.f) {
          const int i = pi[h], j = pj[h];
          const float c = pc[h], sn = ps[h];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int k = lane + 32 * u;
            const float x = S[i * SLD + k], y = S[j * SLD + k];
            S[i * SLD + k] = c * x - sn * y;
            S[j * SLD + k] = sn * x + c * y;
            const float a = J[i * PR + k], b = J[j * PR + k];
            J[i * PR + k] = c * a - sn * b;
            J[j * PR + k] = sn * a + c * b;
          }
        }
      }
      __syncthreads();
      // columns of S with the same coefficients
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (ps[h] != 0.f) {
          const int i = pi[h], j = pj[h];
          const float c = pc[h], sn = ps[h];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int k = lane + 32 * u;
            const float x = S[k * SLD + i], y = S[k * SLD + j];
            S[k * SLD + i] = c * x - sn * y;
            S[k * SLD + j] = sn * x + c * y;
          }
        }
      }
      __syncthreads();
    }
    if (inner == 0 && lane == 0 && cnt) atomicAdd(&n_cnt, cnt);
    if (!__syncthreads_or(any_rot)) break;
  }
  __syncthreads();
  if (threadIdx.x == 0 && n_cnt) atomicAdd(&ctl->rot[sweep], n_cnt);
  if (threadIdx.x < PR) lam[threadIdx.x] = S[threadIdx.x * SLD + threadIdx.x];
  __syncthreads();
  if (threadIdx.x < PR) {
    const int k = threadIdx.x;
    const float v = lam[k];
    int r = 0;
    for (int l = 0; l < PR; ++l) r += (lam[l] > v) || (lam[l] == v && l < k);
    rank[k] = r;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < PR * PR; e += blockDim.x) {
    const int k = e / PR, c = e % PR;
    Jg[rank[k] * PR + c] = J[e];
  }
  if (lam_out && threadIdx.x < PR) lam_out[(int64_t)blockIdx.x * PR + rank[threadIdx.x]] = lam[threadIdx.x];
}

// end of a sweep: an odd number of steps left the data in W1 -> copy it back so every sweep ends in W0
__global__ void sweep_copy_kernel(const float* __restrict__ W1, float* __restrict__ W0, int64_t n4, const Ctl* ctl) {
  if (ctl->done) return;
  const float4* src = reinterpret_cast<const float4*>(W1);
  float4* dst = reinterpret_cast<float4*>(W0);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (int64_t)gridDim.x * blockDim.x) dst[e] = src[e];
}

__global__ void sweep_end_kernel(Ctl* ctl, int sweep, int last) {
  if (ctl->done) return;
  ctl->sweeps = sweep + 1;
  if (ctl->rot[sweep] == 0 || last) ctl->done = 1;
}

__global__ void reopen_kernel(Ctl* ctl) { ctl->done = 0; }

// Vt <- 1.5 Vt - 0.5 T   (one Newton-Schulz step; T = (Vt Vt^T) Vt), both with leading dimension ld
__global__ void newton_schulz_axpy_kernel(float* __restrict__ Vt, const float* __restrict__ Tm, int P, int64_t ld) {
  const int64_t total = (int64_t)P * P;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / P, c = e - r * P;
    Vt[r * ld + c] = 1.5f * Vt[r * ld + c] - 0.5f * Tm[r * ld + c];
  }
}

__global__ void wrow_table_kernel(int64_t* wrow, int64_t n, int64_t ld) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) wrow[e] = e * ld;
}

// per row of the final W: lambda_k = <B_k, Vt_k>, keep_k = (eigenvector lives in the un-padded coordinates)
__global__ void row_stats_kernel(const float* __restrict__ W, int p, int P, float* lam_tmp, int* keep) {
  const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= P) return;
  const float* b = W + k * 2 * P; const float* v = b + P;
  float s = 0.f, sup = 0.f;
  for (int c = lane; c < P; c += 32) {
    const float vv = v[c];
    s += b[c] * vv;
    if (c < p) sup += vv * vv;
  }
  s = wsum(s); sup = wsum(sup);
  if (lane == 0) { lam_tmp[k] = s > 0.f ? s : 0.f; keep[k] = sup > 0.5f ? 1 : 0; }
}

// single CTA: exclusive scan of keep -> destination slot
__global__ void compact_scan_kernel(const int* __restrict__ keep, int P, int* dest) {
  __shared__ int part[1024];
  const int per = (P + blockDim.x - 1) / blockDim.x;
  const int lo = threadIdx.x * per, hi = min(P, lo + per);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += keep[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < (int)blockDim.x; ++i) { const int t = part[i]; part[i] = run; run += t; }
  }
  __syncthreads();
  int run = part[threadIdx.x];
  for (int i = lo; i < hi; ++i) { dest[i] = keep[i] ? run : -1; run += keep[i]; }
}

__global__ void gather_out_kernel(const float* __restrict__ W, int p, int P, const float* __restrict__ lam_tmp,
                                  const int* __restrict__ dest, float* Vt_out, float* lam_out) {
  const int k = blockIdx.x;
  const int d = dest[k];
  if (d < 0 || d >= p) return;
  const float* v = W + (int64_t)k * 2 * P + P;
  for (int c = threadIdx.x; c < p; c += blockDim.x) Vt_out[(int64_t)d * p + c] = v[c];
  if (threadIdx.x == 0) lam_out[d] = lam_tmp[k];
}

// optional measurement callbacks (class 0: pair Gram, 1: rotate rows, 2: refresh); begin returns a slot
struct Hooks {
  int (*begin)(int cls, double flops, cudaStream_t st);
  void (*end)(int slot, cudaStream_t st);
};

struct Plan {
  int p, P, nblk, npairs, steps;
  size_t bytes;
  size_t off_W0, off_W1, off_S, off_J, off_rows, off_jrows, off_iota, off_wrow, off_grow, off_ctl, off_lam, off_keep, off_dest;
};

inline Plan make_plan(int p) {
  Plan pl{};
  pl.p = p;
  int nblk = (p + BLK - 1) / BLK;
  if (nblk < 2) nblk = 2;
  if (nblk & 1) ++nblk;
  pl.nblk = nblk; pl.P = nblk * BLK; pl.npairs = nblk / 2; pl.steps = nblk - 1;
  size_t off = 0;
  auto take = [&](size_t bytes) { off = (off + 255) / 256 * 256; size_t o = off; off += bytes; return o; };
  const size_t P = pl.P;
  pl.off_W0 = take(P * 2 * P * 4);
  pl.off_W1 = take(P * 2 * P * 4);
  pl.off_S = take((size_t)pl.npairs * PR * PR * 4);
  pl.off_J = take((size_t)pl.npairs * PR * PR * 4);
  pl.off_rows = take((size_t)pl.steps * pl.npairs * PR * 8);
  pl.off_jrows = take((size_t)pl.npairs * PR * 8);
  pl.off_iota = take(2 * P * 8);
  pl.off_wrow = take(P * 8);
  pl.off_grow = take(P * 8);
  pl.off_ctl = take(sizeof(Ctl));
  pl.off_lam = take(P * 4);
  pl.off_keep = take(P * 4);
  pl.off_dest = take(P * 4);
  pl.bytes = off + 256;
  return pl;
}

// Enqueue the whole solve.  G: p x p (read only), Vt_out: p x p, lam_out: p.  Returns #kernel launches (<0: CUDA error).
//   phase 1: sweeps until no pair of rows has |cos| > 3e-5 (or kMaxSweeps)
//   refresh: Vt <- Newton-Schulz(Vt) (restores orthogonality lost to fp32 rounding), B <- Vt G (from the original G)
//   phase 2: one more sweep on the refreshed state (quadratic convergence: couplings of ~1e-4 drop below fp32 noise)
//   refresh again, lambda_k = <B_k, Vt_k> (Rayleigh quotients), drop the padding rows.
inline int solve(const float* G, float* Vt_out, float* lam_out, int p, void* ws, cudaStream_t st, cudaError_t* err,
                 const Hooks* hooks = nullptr) {
  const Plan pl = make_plan(p);
  char* base = (char*)ws;
  float* W[2] = {(float*)(base + pl.off_W0), (float*)(base + pl.off_W1)};
  float* S = (float*)(base + pl.off_S);
  float* J = (float*)(base + pl.off_J);
  int64_t* rows = (int64_t*)(base + pl.off_rows);
  int64_t* jrows = (int64_t*)(base + pl.off_jrows);
  int64_t* iota = (int64_t*)(base + pl.off_iota);
  int64_t* wrow = (int64_t*)(base + pl.off_wrow);
  int64_t* grow = (int64_t*)(base + pl.off_grow);
  Ctl* ctl = (Ctl*)(base + pl.off_ctl);
  float* lam_tmp = (float*)(base + pl.off_lam);
  int* keep = (int*)(base + pl.off_keep);
  int* dest = (int*)(base + pl.off_dest);
  const int P = pl.P;
  const int64_t ldw = 2 * (int64_t)P;
  int launches = 0;
#define BJ_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { *err = e_; return -1; } } while (0)
  BJ_CK(cudaMemsetAsync(ctl, 0, sizeof(Ctl), st));
  init_kernel<<<148 * 8, 256, 0, st>>>(G, p, P, W[0]); ++launches;
  pair_tables_kernel<<<148 * 2, 256, 0, st>>>(pl.nblk, ldw, rows); ++launches;
  misc_tables_kernel<<<64, 256, 0, st>>>(pl.npairs, 2 * (int64_t)P, jrows, iota); ++launches;
  wrow_table_kernel<<<32, 256, 0, st>>>(wrow, P, ldw); ++launches;
  wrow_table_kernel<<<32, 256, 0, st>>>(grow, p, p); ++launches;
  BJ_CK(cudaGetLastError());
  static bool attr_set = false;
  if (!attr_set) {
    BJ_CK(cudaFuncSetAttribute(small_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMALL_EIG_SMEM));
    attr_set = true;
  }
  const float tol_rot = 2e-6f, tol_count = 3e-5f;

  auto sweep = [&](int sweep_idx, bool last) -> int {
    int cur = 0;
    for (int s = 0; s < pl.steps; ++s) {
      const int64_t* rows_s = rows + (int64_t)s * pl.npairs * PR;
      TcGemmParams g = tc_gemm_defaults();            // 1. pair Grams
      g.A = {W[cur], rows_s, iota, PR, 2};
      g.B = g.A;
      g.M = PR; g.N = PR; g.K = P; g.C = S; g.ldc = PR; g.batch = pl.npairs;
      g.a_roff_batch = PR; g.b_roff_batch = PR; g.c_batch_stride = PR * PR;
      g.skip_flag = &ctl->done;
      { const int slot = hooks ? hooks->begin(0, 2.0 * PR * PR * (double)P * pl.npairs, st) : -1;
        BJ_CK(tc_gemm_launch(g, 1, 128, st)); ++launches;
        if (hooks) hooks->end(slot, st); }
      small_eig_kernel<<<pl.npairs, 1024, SMALL_EIG_SMEM, st>>>(S, J, tol_rot, tol_count, 6, ctl, sweep_idx); ++launches;  // 2.
      TcGemmParams a = tc_gemm_defaults();            // 3. rotate the pair's rows of [B | Vt]
      a.A = {J, jrows, iota, PR, 2};
      a.B = {W[cur], iota, rows_s, 2 * P, 0};
      a.M = PR; a.N = 2 * P; a.K = PR; a.C = W[cur ^ 1]; a.ldc = ldw; a.batch = pl.npairs;
      a.a_roff_batch = PR; a.b_koff_batch = PR; a.c_roff = rows_s;
      a.skip_flag = &ctl->done;
      { const int slot = hooks ? hooks->begin(1, 2.0 * PR * PR * 2.0 * P * pl.npairs, st) : -1;
        BJ_CK(tc_gemm_launch(a, 1, 256, st)); ++launches;
        if (hooks) hooks->end(slot, st); }
      cur ^= 1;
    }
    // steps is odd: bring the data back to W[0]
    sweep_copy_kernel<<<148 * 8, 256, 0, st>>>(W[1], W[0], (int64_t)P * 2 * P / 4, ctl); ++launches;
    sweep_end_kernel<<<1, 1, 0, st>>>(ctl, sweep_idx, last ? 1 : 0); ++launches;
    return 0;
  };

  auto refresh = [&]() -> int {
    float* Vt = W[0] + P;
    // M1 = Vt Vt^T  -> W[1] (B part)
    TcGemmParams m = tc_gemm_defaults();
    m.A = {Vt, wrow, iota, P, 2}; m.B = m.A;
    m.M = P; m.N = P; m.K = P; m.C = W[1]; m.ldc = ldw; m.symmetric = 1;
    { const int slot = hooks ? hooks->begin(2, (double)P * P * P, st) : -1;
      BJ_CK(tc_gemm_launch(m, 1, 256, st)); ++launches;
      if (hooks) hooks->end(slot, st); }
    // T = M1 Vt -> W[1] (Vt part):  C[i,k] = sum_j M1[i][j] Vt[j][k]
    TcGemmParams t = tc_gemm_defaults();
    t.A = {W[1], wrow, iota, P, 2};
    t.B = {Vt, iota, wrow, P, 0};
    t.M = P; t.N = P; t.K = P; t.C = W[1] + P; t.ldc = ldw;
    { const int slot = hooks ? hooks->begin(2, 2.0 * P * P * P, st) : -1;
      BJ_CK(tc_gemm_launch(t, 1, 256, st)); ++launches;
      if (hooks) hooks->end(slot, st); }
    newton_schulz_axpy_kernel<<<148 * 8, 256, 0, st>>>(Vt, W[1] + P, P, ldw); ++launches;
    // B = Vt G :  C[i,k] = sum_j Vt[i][j] G[k][j]   (G symmetric, un-padded p x p)
    TcGemmParams b = tc_gemm_defaults();
    b.A = {Vt, wrow, iota, P, (p % 4 == 0) ? 2 : 1};
    b.B = {G, grow, iota, p, (p % 4 == 0 && ((uintptr_t)G & 15) == 0) ? 2 : 1};
    b.M = P; b.N = p; b.K = p; b.C = W[0]; b.ldc = ldw;
    { const int slot = hooks ? hooks->begin(2, 2.0 * P * p * p, st) : -1;
      BJ_CK(tc_gemm_launch(b, 1, p > 128 ? 256 : 128, st)); ++launches;
      if (hooks) hooks->end(slot, st); }
    return 0;
  };

  for (int sw = 0; sw < kMaxSweeps - 1; ++sw)
    if (sweep(sw, sw == kMaxSweeps - 2) < 0) return -1;
  if (refresh() < 0) return -1;
  reopen_kernel<<<1, 1, 0, st>>>(ctl); ++launches;
  if (sweep(kMaxSweeps - 1, true) < 0) return -1;
  if (refresh() < 0) return -1;
  row_stats_kernel<<<(P * 32 + 255) / 256, 256, 0, st>>>(W[0], p, P, lam_tmp, keep); ++launches;
  compact_scan_kernel<<<1, 1024, 0, st>>>(keep, P, dest); ++launches;
  gather_out_kernel<<<P, 256, 0, st>>>(W[0], p, P, lam_tmp, dest, Vt_out, lam_out); ++launches;
  BJ_CK(cudaGetLastError());
  if (getenv("MVB_DEBUG_EIG")) {
    Ctl h;
    BJ_CK(cudaStreamSynchronize(st));
    BJ_CK(cudaMemcpy(&h, ctl, sizeof(Ctl), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[mvb eig] p=%d P=%d sweeps=%u rot:", p, P, h.sweeps);
    for (int i = 0; i < kMaxSweeps; ++i) fprintf(stderr, " %u", h.rot[i]);
    fprintf(stderr, "\n");
  }
#undef BJ_CK
  return launches;
}

}  // namespace bj
}  // namespace mvb
