// Standalone check of the split-TF32 tcgen05 contraction core (mvpy_b200/csrc/tc_gemm.cuh)
// against an fp64 host reference.  Build:  make -C tools tc_gemm_test ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cstdint>
#include "../mvpy_b200/csrc/tc_gemm.cuh"

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
  int lanes_along_k;
};

static double run_case(const char* name, const HostOp& A, const HostOp& B, int M, int N, int K, int bn,
                       int splits, int symmetric, int planes = 0) {
  float* dA = to_dev(A.base);
  int64_t* dAr = to_dev(A.roff);
  int64_t* dAk = to_dev(A.koff);
  float* dB = symmetric ? dA : to_dev(B.base);
  int64_t* dBr = to_dev(B.roff);
  int64_t* dBk = to_dev(B.koff);
  float* dC;
  size_t csz = (size_t)M * N;
  CK(cudaMalloc(&dC, csz * splits * sizeof(float)));
  CK(cudaMemset(dC, 0xFF, csz * splits * sizeof(float)));  // NaN fill: unwritten cells show up
  mvb::TcGemmParams p = mvb::tc_gemm_defaults();
  p.A = {dA, dAr, dAk, M, A.lanes_along_k};
  p.B = {dB, dBr, dBk, N, B.lanes_along_k};
  p.M = M; p.N = N; p.K = K; p.C = dC; p.ldc = N; p.c_split_stride = (int64_t)csz; p.symmetric = symmetric; p.dbg = 0;
  uint8_t* dP = nullptr;
  if (planes) {          // B handed over as pre-split planes (TMA path)
    CK(cudaMalloc(&dP, mvb::tc_planes_bytes(N, K)));
    CK(cudaMemset(dP, 0xFF, mvb::tc_planes_bytes(N, K)));
    CK(mvb::tc_presplit_launch(p.B, K, dP, 0));
    p.b_planes = dP; p.b_planes_nkb = (K + 31) / 32;
  }
  CK(mvb::tc_gemm_launch(p, splits, bn, 0));
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
  printf("%-28s M=%d N=%d K=%d bn=%d splits=%d sym=%d  max_abs_err=%.3e  max|ref|=%.3e  rel=%.3e  nan=%ld  %s\n",
         name, M, N, K, bn, splits, symmetric, max_err, max_ref, rel, bad,
         (rel < 1.5e-6 && bad == 0) ? "OK" : "FAIL");
  cudaFree(dA); cudaFree(dAr); cudaFree(dAk); if (!symmetric) cudaFree(dB); cudaFree(dBr); cudaFree(dBk); cudaFree(dC);
  if (dP) cudaFree(dP);
  return (rel < 1.5e-6 && bad == 0) ? 0 : 1;
}

static HostOp dense_op(int rows, int K, int lanes_along_k) {
  HostOp o;
  o.lanes_along_k = lanes_along_k;
  o.base.resize((size_t)rows * K);
  for (auto& v : o.base) v = frand();
  o.roff.resize(rows);
  o.koff.resize(K);
  if (lanes_along_k) {  // row-major [rows][K]
    for (int r = 0; r < rows; ++r) o.roff[r] = (int64_t)r * K;
    for (int k = 0; k < K; ++k) o.koff[k] = k;
  } else {  // stored [K][rows]
    for (int r = 0; r < rows; ++r) o.roff[r] = r;
    for (int k = 0; k < K; ++k) o.koff[k] = (int64_t)k * rows;
  }
  return o;
}

// implicit lag-expanded operand, columns (c,w) as tile rows, (trial,t) as contraction
static HostOp toeplitz_cols(int trials, int C, int T, int W, const std::vector<float>& xp, int lanes_along_k) {
  HostOp o;
  o.lanes_along_k = lanes_along_k;
  const int Tp = T + W - 1;
  o.base = xp;
  for (int c = 0; c < C; ++c)
    for (int w = 0; w < W; ++w) o.roff.push_back((int64_t)c * Tp + w);
  for (int n = 0; n < trials; ++n)
    for (int t = 0; t < T; ++t) o.koff.push_back((int64_t)n * C * Tp + t);
  return o;
}

int main(int argc, char** argv) {
  int fails = 0;
  srand(1);
  {
    HostOp A = dense_op(300, 1000, 1), B = dense_op(520, 1000, 1);
    fails += run_case("dense KK bn256", A, B, 300, 520, 1000, 256, 1, 0);
    fails += run_case("dense KK bn128 split3", A, B, 300, 520, 1000, 128, 3, 0);
    A.lanes_along_k = 2; B.lanes_along_k = 2;
    fails += run_case("dense vec4 bn256", A, B, 300, 520, 1000, 256, 1, 0);
    fails += run_case("dense vec4 bn128 split2", A, B, 300, 520, 1000, 128, 2, 0);
    fails += run_case("dense vec4 x planes", A, B, 300, 520, 1000, 256, 1, 0, 1);
    fails += run_case("dense vec4 x planes split3", A, B, 300, 520, 1000, 256, 3, 0, 1);
    B.lanes_along_k = 1;
    fails += run_case("dense vec4 x planes(scalar)", A, B, 300, 520, 1000, 256, 2, 0, 1);
  }
  {
    HostOp A = dense_op(130, 77, 0), B = dense_op(64, 77, 1);
    fails += run_case("dense RK x planes", A, B, 130, 64, 77, 256, 1, 0, 1);
  }
  {
    HostOp A = dense_op(130, 77, 0), B = dense_op(64, 77, 1);
    fails += run_case("dense RK bn128", A, B, 130, 64, 77, 128, 1, 0);
    fails += run_case("dense RK bn256 split2", A, B, 130, 64, 77, 256, 2, 0);
  }
  {
    int trials = 3, C = 2, T = 150, W = 201, Tp = T + W - 1;
    std::vector<float> xp((size_t)trials * C * Tp);
    for (auto& v : xp) v = frand();
    HostOp G1 = toeplitz_cols(trials, C, T, W, xp, 1);
    HostOp G0 = toeplitz_cols(trials, C, T, W, xp, 0);
    int p = C * W, K = trials * T;
    fails += run_case("toeplitz gram sym bn256", G1, G1, p, p, K, 256, 1, 1);
    fails += run_case("toeplitz gram sym bn128 s2", G0, G0, p, p, K, 128, 2, 1);
    fails += run_case("toeplitz gram mixed", G1, G0, p, p, K, 256, 1, 0);
    fails += run_case("toeplitz gram sym x planes", G0, G0, p, p, K, 256, 1, 1, 1);
    fails += run_case("toeplitz gram sym x planes s2", G0, G0, p, p, K, 256, 2, 1, 1);
    HostOp D = dense_op(700, K, 1);      // K = 450 is not a multiple of 4: scalar (mode 1) source
    fails += run_case("toeplitz x dense planes", G0, D, p, 700, K, 256, 1, 0, 1);
    fails += run_case("toeplitz(k) x dense planes s2", G1, D, p, 700, K, 256, 2, 0, 1);
  }
  if (argc > 1) {  // throughput: dense 8192 x 8192 x 8192
    int M = 8192, N = 8192, K = 8192;
    for (int mode = 0; mode < 3; ++mode) {
      HostOp A = dense_op(M, K, 1), B = dense_op(N, K, mode == 1 ? 0 : 1);
      if (mode == 2) { A.lanes_along_k = 2; B.lanes_along_k = 2; }
      float* dA = to_dev(A.base); int64_t* dAr = to_dev(A.roff); int64_t* dAk = to_dev(A.koff);
      float* dB = to_dev(B.base); int64_t* dBr = to_dev(B.roff); int64_t* dBk = to_dev(B.koff);
      float* dC; CK(cudaMalloc(&dC, (size_t)M * N * 4));
      mvb::TcGemmParams p = mvb::tc_gemm_defaults();
      p.A = {dA, dAr, dAk, M, A.lanes_along_k}; p.B = {dB, dBr, dBk, N, B.lanes_along_k};
      p.M = M; p.N = N; p.K = K; p.C = dC; p.ldc = N; p.c_split_stride = 0; p.symmetric = 0;
      if (mode == 2) {
        uint8_t* dP; CK(cudaMalloc(&dP, mvb::tc_planes_bytes(N, K)));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        CK(mvb::tc_presplit_launch(p.B, K, dP, 0));
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float msp; cudaEventElapsedTime(&msp, e0, e1);
        mvb::TcGemmParams pp = p; pp.b_planes = dP; pp.b_planes_nkb = (K + 31) / 32; pp.dbg = 0;
        for (int rast = 0; rast < 2; ++rast) {
          pp.raster_m_fast = rast;
          CK(mvb::tc_gemm_launch(pp, 1, 256, 0)); CK(cudaDeviceSynchronize());
          cudaEventRecord(e0);
          for (int i = 0; i < 5; ++i) CK(mvb::tc_gemm_launch(pp, 1, 256, 0));
          cudaEventRecord(e1); CK(cudaDeviceSynchronize());
          float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
          printf("perf PLANES mode=2 A x pre-split B (raster_m_fast=%d): %.3f ms  %.1f TFLOP/s (fp32-equivalent); presplit %.3f ms\n", rast, ms,
                 2.0 * M * N * K / ms * 1e-9, msp);
        }
        cudaFree(dP);
      }
      for (int dbg = 0; dbg < 3; ++dbg)
      for (int bn : {256, 128}) {
        p.dbg = dbg;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        CK(mvb::tc_gemm_launch(p, 1, bn, 0));
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) CK(mvb::tc_gemm_launch(p, 1, bn, 0));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("perf dbg=%d mode=%d bn=%d: %.3f ms  %.1f TFLOP/s (fp32-equivalent), %.1f issued TF32 TFLOP/s\n", dbg, mode, bn,
               ms, 2.0 * M * N * K / ms * 1e-9, 6.0 * M * N * K / ms * 1e-9);
      }
      cudaFree(dA); cudaFree(dAr); cudaFree(dAk); cudaFree(dB); cudaFree(dBr); cudaFree(dBk); cudaFree(dC);
    }
  }
  printf("tc_gemm_test: %s (%d failing cases)\n", fails ? "FAILED" : "PASSED", fails);
  return fails ? 1 : 0;
}
