// Host<->device copy bandwidth with page-locked memory: H2D alone, D2H alone, both directions at once.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pcie pcie.cu
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  const size_t n = 2ull << 30;
  char *h1, *h2, *d1, *d2;
  cudaHostAlloc(&h1, n, cudaHostAllocPortable); cudaHostAlloc(&h2, n, cudaHostAllocPortable);
  memset(h1, 1, n); memset(h2, 2, n);
  cudaMalloc(&d1, n); cudaMalloc(&d2, n);
  cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
  for (int rep = 0; rep < 2; ++rep) {
    double t = now(); cudaMemcpyAsync(d1, h1, n, cudaMemcpyHostToDevice, s1); cudaStreamSynchronize(s1);
    printf("H2D alone      %6.1f GB/s\n", n / (now() - t) / 1e9);
    t = now(); cudaMemcpyAsync(h2, d2, n, cudaMemcpyDeviceToHost, s2); cudaStreamSynchronize(s2);
    printf("D2H alone      %6.1f GB/s\n", n / (now() - t) / 1e9);
    t = now(); cudaMemcpyAsync(d1, h1, n, cudaMemcpyHostToDevice, s1); cudaMemcpyAsync(h2, d2, n, cudaMemcpyDeviceToHost, s2);
    cudaStreamSynchronize(s1); cudaStreamSynchronize(s2);
    printf("H2D + D2H      %6.1f GB/s total\n", 2.0 * n / (now() - t) / 1e9);
    // chunked: 64 copies of 8 MB each way
    t = now();
    for (int i = 0; i < 256; ++i) { cudaMemcpyAsync(d1 + i * (8ull << 20), h1 + i * (8ull << 20), 8ull << 20, cudaMemcpyHostToDevice, s1);
                                    cudaMemcpyAsync(h2 + i * (8ull << 20), d2 + i * (8ull << 20), 8ull << 20, cudaMemcpyDeviceToHost, s2); }
    cudaStreamSynchronize(s1); cudaStreamSynchronize(s2);
    printf("H2D + D2H 8MB  %6.1f GB/s total\n", 2.0 * n / (now() - t) / 1e9);
  }
  double t = now(); void *p; cudaMalloc(&p, 1 << 28); printf("cudaMalloc 256MB %.3f ms\n", (now() - t) * 1e3);
  t = now(); cudaFree(p); printf("cudaFree %.3f ms\n", (now() - t) * 1e3);
  t = now(); char *hp; cudaHostAlloc(&hp, 1ull << 30, cudaHostAllocPortable); printf("cudaHostAlloc 1GB %.1f ms\n", (now() - t) * 1e3);
  return 0;
}
