// Standalone check of the warp-specialised split-TF32 contraction core (mvpy_b200/csrc/tc_gemm2.cuh)
// against an fp64 host reference, plus throughput on the hot-path shapes.
// Build:  make -C tools tc_gemm2_test ;  run on a B200:  tools/tc_gemm2_test [perf]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cstdint>
#include "tc_gemm2.cuh"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

template <class T>
T* to_dev(const std::vector<T>& h) {
  T* d;
  CK(cudaMalloc(&d, h.size() * sizeof(T)));
  CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

struct HostOp {
  std::vector<float> base;
  std::vector<int64_t> roff, koff;
  int major = 0;
  int copies = 1;
};

struct DevOp {
  float *x = nullptr, *hi = nullptr, *lo = nullptr;
  int64_t *roff = nullptr, *koff = nullptr;
  int64_t cs = 0;
  mvb::PlaneOperand op(int rows, int, int copies) const { return {hi, lo, copies > 1 ? cs : 0, roff, koff, rows}; }
  void free_all() { cudaFree(x); cudaFree(hi); cudaFree(lo); cudaFree(roff); cudaFree(koff); }
};

static DevOp upload(const HostOp& h) {
  DevOp d;
  d.x = to_dev(h.base);
  d.roff = to_dev(h.roff);
  d.koff = to_dev(h.koff);
  const int64_t n = (int64_t)h.base.size();
  d.cs = (n + 3) / 4 * 4 + 4;
  CK(cudaMalloc(&d.hi, (size_t)h.copies * d.cs * 4));
  CK(cudaMalloc(&d.lo, (size_t)h.copies * d.cs * 4));
  mvb::split_planes_kernel<<<148 * 4, 256>>>(d.x, n, h.copies, d.cs, d.hi, d.lo);
  CK(cudaGetLastError());
  return d;
}

static int run_case(const char* name, const HostOp& A, const HostOp& B, int M, int N, int K, int bn, int splits, int symmetric) {
  DevOp dA = upload(A), dB = symmetric ? dA : upload(B);
  float* dC;
  size_t csz = (size_t)M * N;
  CK(cudaMalloc(&dC, csz * splits * sizeof(float)));
  CK(cudaMemset(dC, 0xFF, csz * splits * sizeof(float)));  // NaN fill: unwritten cells show up
  mvb::TcGemm2Params p = mvb::tc_gemm2_defaults();
  p.A = dA.op(M, A.major, A.copies);
  p.B = dB.op(N, B.major, B.copies);
  p.M = M; p.N = N; p.K = K; p.C = dC; p.ldc = N; p.c_split_stride = (int64_t)csz; p.symmetric = symmetric;
  CK(mvb::tc_gemm2_launch(p, splits, bn, 0));
  CK(cudaDeviceSynchronize());
  std::vector<float> C(csz * splits);
  CK(cudaMemcpy(C.data(), dC, C.size() * sizeof(float), cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0;
  long bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k)
        ref += (double)A.base[A.roff[m] + A.koff[k]] * (double)B.base[B.roff[n] + B.koff[k]];
      double got = 0;
      for (int z = 0; z < splits; ++z) got += C[(size_t)z * csz + (size_t)m * N + n];
      double err = fabs(got - ref);
      if (!(err == err)) { bad++; continue; }
      if (err > max_err) max_err = err;
      if (fabs(ref) > max_ref) max_ref = fabs(ref);
    }
  double rel = max_err / (max_ref + 1e-30);
  const bool ok = rel < 1.5e-6 && bad == 0;
  printf("%-30s M=%d N=%d K=%d bn=%d splits=%d sym=%d majors=%d%d copies=%d%d  max_abs_err=%.3e  max|ref|=%.3e  rel=%.3e  nan=%ld  %s\n",
         name, M, N, K, bn, splits, symmetric, A.major, B.major, A.copies, B.copies, max_err, max_ref, rel, bad, ok ? "OK" : "FAIL");
  fflush(stdout);
  dA.free_all();
  if (!symmetric) dB.free_all();
  cudaFree(dC);
  return ok ? 0 : 1;
}

// dense operand; k_contig: stored row-major [rows][ld] (k contiguous), else stored [K][ld] (rows contiguous)
static HostOp dense_op(int rows, int K, bool k_contig, int major, int ld_pad = 0) {
  HostOp o;
  o.major = major;
  const int ld = (k_contig ? K : rows) + ld_pad;
  o.base.resize((size_t)(k_contig ? rows : K) * ld);
  for (auto& v : o.base) v = frand();
  o.roff.resize(rows);
  o.koff.resize(K);
  if (k_contig) {
    for (int r = 0; r < rows; ++r) o.roff[r] = (int64_t)r * ld;
    for (int k = 0; k < K; ++k) o.koff[k] = k;
  } else {
    for (int r = 0; r < rows; ++r) o.roff[r] = r;
    for (int k = 0; k < K; ++k) o.koff[k] = (int64_t)k * ld;
  }
  return o;
}

// implicit lag-expanded operand, columns (c,w) as tile rows, (trial,t) as contraction
static HostOp toeplitz_cols(int trials, int C, int T, int W, const std::vector<float>& xp) {
  HostOp o;
  o.major = 0; o.copies = 4;
  const int Tp = T + W - 1;
  o.base = xp;
  for (int c = 0; c < C; ++c)
    for (int w = 0; w < W; ++w) o.roff.push_back((int64_t)c * Tp + w);
  for (int n = 0; n < trials; ++n)
    for (int t = 0; t < T; ++t) o.koff.push_back((int64_t)n * C * Tp + t);
  return o;
}
// implicit lag-expanded operand, (trial,t) as tile rows, (c,w) as contraction
static HostOp toeplitz_rows(int trials, int C, int T, int W, const std::vector<float>& xp) {
  HostOp o;
  o.major = 0; o.copies = 4;
  const int Tp = T + W - 1;
  o.base = xp;
  for (int n = 0; n < trials; ++n)
    for (int t = 0; t < T; ++t) o.roff.push_back((int64_t)n * C * Tp + t);
  for (int c = 0; c < C; ++c)
    for (int w = 0; w < W; ++w) o.koff.push_back((int64_t)c * Tp + w);
  return o;
}

static void perf_case(const char* name, const HostOp& A, const HostOp& B, int M, int N, int K, int bn, int splits, int symmetric) {
  DevOp dA = upload(A), dB = symmetric ? dA : upload(B);
  float* dC;
  CK(cudaMalloc(&dC, (size_t)M * N * splits * 4));
  mvb::TcGemm2Params p = mvb::tc_gemm2_defaults();
  p.A = dA.op(M, A.major, A.copies);
  p.B = dB.op(N, B.major, B.copies);
  p.M = M; p.N = N; p.K = K; p.C = dC; p.ldc = N; p.c_split_stride = (int64_t)M * N; p.symmetric = symmetric;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  CK(mvb::tc_gemm2_launch(p, splits, bn, 0));
  CK(cudaDeviceSynchronize());
  const int reps = 5;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) CK(mvb::tc_gemm2_launch(p, splits, bn, 0));
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  double fl = 2.0 * M * (double)N * K * (symmetric ? 0.5 : 1.0);
  printf("perf %-34s M=%d N=%d K=%d bn=%d splits=%d sym=%d majors=%d%d: %.3f ms  %.1f TFLOP/s (fp32-equivalent), %.1f issued TF32 TFLOP/s\n",
         name, M, N, K, bn, splits, symmetric, A.major, B.major, ms, fl / ms * 1e-9, 3 * fl / ms * 1e-9);
  fflush(stdout);
  dA.free_all();
  if (!symmetric) dB.free_all();
  cudaFree(dC);
}

int main(int argc, char** argv) {
  int fails = 0;
  srand(1);
  if (argc > 1 && argv[1][0] == 'o') {   // "one": a single dense launch pair, for ncu captures
    int M = 4096, N = 4096, K = 8192;
    HostOp A = dense_op(M, K, true, 0), B = dense_op(N, K, true, 0);
    perf_case("dense 4096x4096x8192 KxK", A, B, M, N, K, 256, 1, 0);
    return 0;
  }
  {
    HostOp A = dense_op(300, 1000, true, 0), B = dense_op(520, 1000, true, 0);
    fails += run_case("dense KK bn256", A, B, 300, 520, 1000, 256, 1, 0);
    fails += run_case("dense KK bn128 split3", A, B, 300, 520, 1000, 128, 3, 0);
    fails += run_case("dense KK bn256 split2", A, B, 300, 520, 1000, 256, 2, 0);
  }
  {
    // slow paths: K tail, unaligned leading dimension, chunk direction that is not the contiguous one
    HostOp A = dense_op(130, 77, false, 0), B = dense_op(64, 77, true, 0);
    fails += run_case("dense strided-K, K tail", A, B, 130, 64, 77, 128, 1, 0);
    HostOp A3 = dense_op(130, 77, true, 0, 1), B3 = dense_op(200, 77, false, 0, 3);
    fails += run_case("dense unaligned ld", A3, B3, 130, 200, 77, 256, 2, 0);
    HostOp A4 = dense_op(130, 2103, true, 0, 1), B4 = dense_op(200, 2103, true, 0);
    fails += run_case("dense unaligned ld long K", A4, B4, 130, 200, 2103, 256, 1, 0);
  }
  {
    int trials = 3, C = 2, T = 152, W = 201, Tp = T + W - 1;
    std::vector<float> xp((size_t)trials * C * Tp);
    for (auto& v : xp) v = frand();
    HostOp G = toeplitz_cols(trials, C, T, W, xp);
    int p = C * W, K = trials * T;
    fails += run_case("toeplitz gram sym bn256", G, G, p, p, K, 256, 1, 1);
    fails += run_case("toeplitz gram sym bn128 s2", G, G, p, p, K, 128, 2, 1);
    fails += run_case("toeplitz gram nonsym", G, G, p, p, K, 256, 1, 0);
    HostOp R = toeplitz_rows(trials, C, T, W, xp);
    HostOp Q = dense_op(300, p, true, 0);
    fails += run_case("toeplitz rows x dense", R, Q, K, 300, p, 256, 1, 0);
    fails += run_case("toeplitz rows x dense bn128", R, Q, K, 300, p, 128, 1, 0);
  }
  printf("tc_gemm2_test: %s (%d failing cases)\n", fails ? "FAILED" : "PASSED", fails);
  if (argc > 1 && !fails) {
    {
      int M = 8192, N = 8192, K = 8192;
      HostOp A = dense_op(M, K, true, 0), B = dense_op(N, K, true, 0);
      perf_case("dense 8192^3 KxK", A, B, M, N, K, 256, 1, 0);
      perf_case("dense 8192^3 KxK bn128", A, B, M, N, K, 128, 1, 0);
      HostOp Bs = dense_op(N, K, false, 0);
      perf_case("dense 8192^3 K x strided-k (4 B copies)", A, Bs, M, N, K, 256, 1, 0);
    }
    {
      int P = 12928;
      HostOp G = dense_op(P, P, true, 0), Q = dense_op(128, P, true, 0);
      perf_case("lanczos W = Qj G", Q, G, 128, P, P, 256, 3, 0);
      perf_case("lanczos W = Qj G bn128", Q, G, 128, P, P, 128, 2, 0);
      HostOp Qc = dense_op(P, P / 2, true, 0), H = dense_op(128, P / 2, true, 0);
      perf_case("lanczos W2 = H Q (Q^T copy)", H, Qc, 128, P, P / 2, 256, 3, 0);
      HostOp Qr = dense_op(P / 2, P, true, 0);
      perf_case("lanczos H = W Q^T", Q, Qr, 128, P / 2, P, 256, 6, 0);
    }
    {
      int trials = 96, C = 64, T = 400, W = 201, Tp = T + W - 1;
      std::vector<float> xp((size_t)trials * C * Tp);
      for (auto& v : xp) v = frand();
      HostOp G = toeplitz_cols(trials, C, T, W, xp);
      int p = C * W, K = trials * T;
      perf_case("gram cfg2 (implicit Toeplitz)", G, G, p, p, K, 256, 1, 1);
      HostOp R = toeplitz_rows(trials, C, T, W, xp);
      HostOp Q = dense_op(4096, p, true, 0);
      perf_case("Z = X Q^T (4096 basis rows)", R, Q, K, 4096, p, 256, 1, 0);
    }
  }
  return fails ? 1 : 0;
}
