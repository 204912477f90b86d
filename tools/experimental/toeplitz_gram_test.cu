// Harness of mvpy_b200/csrc/toeplitz_gram.cuh: correctness against a double-precision host Gram of the explicit lag expansion,
// and throughput at the BASELINE configs[1] shape (96 training trials x 64 features x 400 samples, 201 lags).
//   tools/toeplitz_gram_test [perf]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../mvpy_b200/csrc/toeplitz_gram.cuh"
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)
using namespace mvb;
static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

struct Case { int N, C, T, W; std::vector<int> trials; };

static double run(const Case& cs, bool check, int reps) {
  const int Tp = cs.T + cs.W - 1;
  const int64_t total = (int64_t)cs.N * cs.C * Tp;
  std::vector<float> x(total);
  for (auto& v : x) v = frand();
  const int p = cs.C * cs.W, nt = (int)cs.trials.size();
  float *dx, *dpack, *dG; int32_t* dtidx;
  const int Tpp = tg::pack_pitch(Tp);
  CK(cudaMalloc(&dx, total * 4)); CK(cudaMalloc(&dpack, tg::pack_bytes(nt, cs.C, Tp))); CK(cudaMalloc(&dG, (size_t)p * p * 4));
  CK(cudaMemcpy(dx, x.data(), total * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dG, 0xFF, (size_t)p * p * 4));
  const int nblk_f = (cs.W + 31) / 32, n_vblk = cs.C * nblk_f;
  CK(cudaMalloc(&dtidx, nt * 4));
  CK(cudaMemcpy(dtidx, cs.trials.data(), nt * 4, cudaMemcpyHostToDevice));
  tg::Params prm{};
  prm.pack = dpack; prm.L = (int64_t)nt * cs.C * Tpp; prm.n_trials = nt; prm.n_feat = cs.C; prm.Tpp = Tpp; prm.n_vblk = n_vblk; prm.nblk_f = nblk_f;
  prm.W = cs.W; prm.T = cs.T; prm.G = dG; prm.ldg = p;
  tg::chunk_plan(cs.T, prm.n_chunks, prm.chunk_len);
  cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  double best = 1e30, best_pack = 1e30;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    tg::pack_kernel<<<148 * 8, 256>>>(dx, dtidx, nt, (int64_t)cs.C * Tp, nullptr, cs.C, Tp, Tp, Tpp, dpack);
    cudaEventRecord(e1);
    CK(tg::launch(prm, 0));
    cudaEventRecord(e2);
    CK(cudaDeviceSynchronize());
    float a, b; cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2);
    if (b < best) best = b;
    if (a < best_pack) best_pack = a;
  }
  const double n = (double)nt * cs.T;
  printf("toeplitz gram N=%d C=%d T=%d W=%d (p=%d, n=%.0f, chunks %d x %d): %.3f ms (pack %.3f ms)  %.1f TFLOP/s algorithmic n p (p+1)\n",
         nt, cs.C, cs.T, cs.W, p, n, prm.n_chunks, prm.chunk_len, best, best_pack, n * p * (p + 1.0) / best * 1e-9);
  if (check) {
    std::vector<float> G((size_t)p * p);
    CK(cudaMemcpy(G.data(), dG, G.size() * 4, cudaMemcpyDeviceToHost));
    // explicit design rows in double
    const int nrow = nt * cs.T;
    std::vector<double> D((size_t)nrow * p);
    for (int i = 0; i < nt; ++i) for (int t = 0; t < cs.T; ++t) for (int c = 0; c < cs.C; ++c) for (int w = 0; w < cs.W; ++w)
      D[((size_t)i * cs.T + t) * p + c * cs.W + w] = x[((int64_t)cs.trials[i] * cs.C + c) * Tp + t + w];
    double max_err = 0, max_ref = 0; long bad = 0;
    std::vector<double> colr(nrow);
    for (int a = 0; a < p; ++a) {
      for (int i = 0; i < nrow; ++i) colr[i] = D[(size_t)i * p + a];
      for (int b = 0; b < p; ++b) {
        double ref = 0;
        for (int i = 0; i < nrow; ++i) ref += colr[i] * D[(size_t)i * p + b];
        const double got = G[(size_t)a * p + b], err = fabs(got - ref);
        if (!(err == err)) { ++bad; continue; }
        if (err > max_err) max_err = err;
        if (fabs(ref) > max_ref) max_ref = fabs(ref);
      }
    }
    const bool ok = bad == 0 && max_err / max_ref < 2e-6;
    printf("   check: max_abs_err=%.3e max|ref|=%.3e rel=%.3e nan=%ld %s\n", max_err, max_ref, max_err / max_ref, bad, ok ? "OK" : "FAIL");
    if (!ok) return -1;
  }
  cudaFree(dx); cudaFree(dpack); cudaFree(dG); cudaFree(dtidx);
  return best;
}

int main(int argc, char** argv) {
  srand(3);
  bool ok = true;
  ok &= run({3, 3, 64, 37, {2, 0}}, true, 1) >= 0;          // 6 virtual blocks: two M tiles, one N tile, ragged last block
  ok &= run({4, 5, 400, 201, {0, 1, 3}}, true, 1) >= 0;     // cfg2 geometry per feature (7 blocks, 9 valid rows in the last), 3 chunks per trial
  ok &= run({2, 9, 152, 64, {1}}, true, 1) >= 0;            // W multiple of 32, two unequal chunks (80 + 72)
  printf("toeplitz_gram_test: %s\n", ok ? "PASSED" : "FAILED");
  if (!ok) return 1;
  if (argc > 1) {
    std::vector<int> tr(96); for (int i = 0; i < 96; ++i) tr[i] = i + (i >= 24 ? 24 : 0);     // a training fold of 120 trials
    run({120, 64, 400, 201, tr}, false, 5);
  }
  return 0;
}
