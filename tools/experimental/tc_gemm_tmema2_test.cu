// EXPERIMENT: A operand in tensor memory + B planes by TMA (tc_gemm_tmema2.cuh) vs fp64 reference + throughput.  nvcc -arch sm_100a ...
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "tc_gemm_tmema2.cuh"
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)
static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }
int main(int argc, char** argv) {
  srand(1);
  {
    const int M = 512, N = 768, K = 1056;   // 33 stages: both halves end in a ragged chunk
    std::vector<float> A((size_t)M * K), B((size_t)N * K), C((size_t)M * N);
    for (auto& v : A) v = frand();
    for (auto& v : B) v = frand();
    float *dA, *dB, *dC;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dC, C.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0xFF, C.size() * 4));
    uint8_t* dP; CK(cudaMalloc(&dP, mvb::tc_planes_bytes(N, K)));
    { std::vector<int64_t> ro(N), ko(K); for (int i = 0; i < N; ++i) ro[i] = (int64_t)i * K; for (int i = 0; i < K; ++i) ko[i] = i;
      int64_t *dro, *dko; CK(cudaMalloc(&dro, N * 8)); CK(cudaMalloc(&dko, K * 8)); CK(cudaMemcpy(dro, ro.data(), N * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dko, ko.data(), K * 8, cudaMemcpyHostToDevice));
      mvb::GatherOperand op{dB, dro, dko, N, 2}; CK(mvb::tc_presplit_launch(op, K, dP, 0)); CK(cudaDeviceSynchronize()); }
    CK(mvb::tmema2::tmema2_launch(dA, dP, dC, M, N, K, 0));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0, max_ref = 0; long bad = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
      double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
      double err = fabs(C[(size_t)m * N + n] - ref);
      if (!(err == err)) { ++bad; continue; }
      if (err > max_err) max_err = err; if (fabs(ref) > max_ref) max_ref = fabs(ref);
    }
    printf("tmem-A (staggered fold) gemm %dx%dx%d: max_abs_err=%.3e max|ref|=%.3e rel=%.3e nan=%ld %s\n", M, N, K, max_err, max_ref, max_err / max_ref, bad,
           (max_err / max_ref < 1.5e-6 && bad == 0) ? "OK" : "FAIL");
    fflush(stdout);
    if (!(max_err / max_ref < 1.5e-6 && bad == 0)) return 1;
  }
  if (argc > 1) {
    const int M = 8192, N = 8192, K = 8192;
    std::vector<float> A((size_t)M * K);
    for (auto& v : A) v = frand();
    float *dA, *dB, *dC;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, A.size() * 4)); CK(cudaMalloc(&dC, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    uint8_t* dP; CK(cudaMalloc(&dP, mvb::tc_planes_bytes(N, K)));
    { std::vector<int64_t> ro(N), ko(K); for (int i = 0; i < N; ++i) ro[i] = (int64_t)i * K; for (int i = 0; i < K; ++i) ko[i] = i;
      int64_t *dro, *dko; CK(cudaMalloc(&dro, N * 8)); CK(cudaMalloc(&dko, K * 8)); CK(cudaMemcpy(dro, ro.data(), N * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dko, ko.data(), K * 8, cudaMemcpyHostToDevice));
      mvb::GatherOperand op{dB, dro, dko, N, 2}; CK(mvb::tc_presplit_launch(op, K, dP, 0)); CK(cudaDeviceSynchronize()); }
    CK(mvb::tmema2::tmema2_launch(dA, dP, dC, M, N, K, 0)); CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) CK(mvb::tmema2::tmema2_launch(dA, dP, dC, M, N, K, 0));
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    printf("perf tmem-A (staggered fold) gemm 8192^3: %.3f ms  %.1f TFLOP/s (fp32-equivalent)\n", ms, 2.0 * M * N * K / ms * 1e-9);
  }
  return 0;
}
