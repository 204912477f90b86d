// Micro-benchmark (written at the end of r01, not yet run): how fast can one CTA per SM stream HBM through cp.async.bulk, as a
// function of the copy size and the number of copies in flight?  Decides the ring shape for HBM-streamed operands (the skinny
// Lanczos products measured ~16 B/clk per SM with two 64 KB copies in flight; the L2-fed products 21 B/clk).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_stream_bench tools/experimental/tma_stream_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../mvpy_b200/csrc/tc_gemm.cuh"
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)
using namespace mvb::tc;

// every CTA streams bytes_per_cta contiguous bytes starting at src + blockIdx.x * bytes_per_cta through `depth` slots of `slot` bytes
__global__ void __launch_bounds__(128, 1) stream_kernel(const uint8_t* __restrict__ src, size_t bytes_per_cta, int slot, int depth,
                                                        unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)slot * depth);
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint8_t* p = src + (size_t)blockIdx.x * bytes_per_cta;
    const int n = (int)(bytes_per_cta / slot);
    for (int i = 0; i < depth && i < n; ++i) bulk_load(smem_u32(smem + (size_t)i * slot), p + (size_t)i * slot, slot, smem_u32(&bars[i]));
    unsigned long long acc = 0;
    for (int i = 0; i < n; ++i) {
      const int s = i % depth;
      mbar_wait(smem_u32(&bars[s]), (uint32_t)((i / depth) & 1));
      acc += *reinterpret_cast<volatile unsigned int*>(smem + (size_t)s * slot);          // consume something
      if (i + depth < n) bulk_load(smem_u32(smem + (size_t)s * slot), p + (size_t)(i + depth) * slot, slot, smem_u32(&bars[s]));
    }
    if (acc == 0x123456789abcdefull) *sink = acc;
  }
}

int main() {
  const size_t total = (size_t)4 << 30;                  // 4 GB: every launch reads from HBM (L2 is 126 MB)
  uint8_t* d; unsigned long long* sink;
  CK(cudaMalloc(&d, total)); CK(cudaMemset(d, 1, total)); CK(cudaMalloc(&sink, 8));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("%d SMs, one CTA each, %zu MB per CTA per launch\n", sms, total / sms >> 20);
  for (int slot : {8192, 16384, 32768, 65536}) {
    for (int depth : {1, 2, 3, 4, 6, 8, 12, 16, 24}) {
      if ((size_t)slot * depth > 192 * 1024) continue;
      const size_t per = (total / sms) / ((size_t)slot * depth) * ((size_t)slot * depth);
      const size_t smem_bytes = (size_t)slot * depth + 256;
      stream_kernel<<<sms, 128, smem_bytes>>>(d, per, slot, depth, sink); CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      for (int r = 0; r < 3; ++r) stream_kernel<<<sms, 128, smem_bytes>>>(d, per, slot, depth, sink);
      cudaEventRecord(e1); CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
      const double gbs = (double)per * sms / ms * 1e-6;
      printf("slot %6d B x depth %2d (%3zu KB in flight per SM): %7.3f ms  %7.1f GB/s  %5.1f B/clk/SM at 1.965 GHz\n", slot, depth,
             (size_t)slot * depth >> 10, ms, gbs, gbs * 1e9 / sms / 1.965e9);
    }
  }
  return 0;
}
