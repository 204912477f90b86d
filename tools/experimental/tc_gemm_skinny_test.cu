// EXPERIMENT: turned-around skinny product (tc_gemm_skinny.cuh) vs fp64 reference + throughput at the cfg2 Lanczos size.
//   make -C tools tc_gemm_skinny_test && tools/tc_gemm_skinny_test perf
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "tc_gemm_skinny.cuh"
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)
static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

static int check(int N, int K, int splits) {
  std::vector<float> Q((size_t)128 * K), G((size_t)N * K), W((size_t)splits * 128 * N);
  for (auto& v : Q) v = frand();
  for (auto& v : G) v = frand();
  float *dQ, *dG, *dW; uint8_t* dP;
  CK(cudaMalloc(&dQ, Q.size() * 4)); CK(cudaMalloc(&dG, G.size() * 4)); CK(cudaMalloc(&dW, W.size() * 4));
  CK(cudaMalloc(&dP, mvb::skinny::planes128_bytes(K)));
  CK(cudaMemcpy(dQ, Q.data(), Q.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dG, G.data(), G.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dW, 0xFF, W.size() * 4)); CK(cudaMemset(dP, 0xFF, mvb::skinny::planes128_bytes(K)));
  CK(mvb::skinny::presplit128_launch(dQ, K, K, dP, 0));
  CK(mvb::skinny::skinny_launch(dG, K, dP, dW, N, (int64_t)128 * N, N, K, splits, 0));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(W.data(), dW, W.size() * 4, cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0; long bad = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
    double ref = 0; for (int k = 0; k < K; ++k) ref += (double)Q[(size_t)m * K + k] * G[(size_t)n * K + k];
    double got = 0; for (int z = 0; z < splits; ++z) got += W[(size_t)z * 128 * N + (size_t)m * N + n];
    double err = fabs(got - ref);
    if (!(err == err)) { ++bad; continue; }
    if (err > max_err) max_err = err; if (fabs(ref) > max_ref) max_ref = fabs(ref);
  }
  const bool ok = max_err / max_ref < 1.5e-6 && bad == 0;
  printf("skinny 128x%dx%d splits=%d: max_abs_err=%.3e max|ref|=%.3e rel=%.3e nan=%ld %s\n", N, K, splits, max_err, max_ref, max_err / max_ref, bad, ok ? "OK" : "FAIL");
  fflush(stdout);
  cudaFree(dQ); cudaFree(dG); cudaFree(dW); cudaFree(dP);
  return ok ? 0 : 1;
}

int main(int argc, char** argv) {
  srand(1);
  int fails = 0;
  fails += check(512, 1024, 1);
  fails += check(384, 416, 2);      // 13 stages: ragged last chunk, uneven split
  fails += check(256, 96, 1);       // fewer stages than the ring depth
  fails += check(640, 2048, 3);
  if (fails) { printf("tc_gemm_skinny_test: FAILED (%d)\n", fails); return 1; }
  if (argc > 1) {
    const int N = 12928, K = 12928;                 // cfg2: P = 12 928
    std::vector<float> G((size_t)N * K);
    for (size_t i = 0; i < G.size(); ++i) G[i] = (float)((i * 2654435761u) >> 8 & 0xFFFF) / 65536.f - 0.5f;
    float *dQ, *dG, *dW; uint8_t* dP;
    CK(cudaMalloc(&dQ, (size_t)128 * K * 4)); CK(cudaMalloc(&dG, G.size() * 4)); CK(cudaMalloc(&dW, (size_t)8 * 128 * N * 4));
    CK(cudaMalloc(&dP, mvb::skinny::planes128_bytes(K)));
    CK(cudaMemcpy(dG, G.data(), G.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dQ, G.data(), (size_t)128 * K * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int splits : {1, 2, 3, 4, 6, 8}) {
      CK(mvb::skinny::presplit128_launch(dQ, K, K, dP, 0));
      CK(mvb::skinny::skinny_launch(dG, K, dP, dW, N, (int64_t)128 * N, N, K, splits, 0)); CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      for (int i = 0; i < 10; ++i) {               // G (0.67 GB) exceeds L2: every launch streams it from HBM
        CK(mvb::skinny::presplit128_launch(dQ, K, K, dP, 0));
        CK(mvb::skinny::skinny_launch(dG, K, dP, dW, N, (int64_t)128 * N, N, K, splits, 0));
      }
      cudaEventRecord(e1); CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
      printf("perf skinny 128 x %d x %d splits=%d (presplit of the block included): %.3f ms  %.1f TFLOP/s  %.2f TB/s of G\n", N, K, splits, ms,
             2.0 * 128 * N * K / ms * 1e-9, (double)N * K * 4 / ms * 1e-9);
    }
  }
  printf("tc_gemm_skinny_test: PASSED\n");
  return 0;
}
