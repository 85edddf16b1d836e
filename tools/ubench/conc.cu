// Do two kernels on two non-blocking streams overlap on this box?  (mode 3 of the sweep relies on it.)
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o conc conc.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void spin(long long cycles, int *sink) {
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {}
  if (sink && threadIdx.x == 12345) *sink = 1;
}
static float run(int gridA, int gridB, bool ev_between, int smem) {
  cudaStream_t a, b;
  cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
  cudaEvent_t t0, t1, fork, ea, eb;
  cudaEventCreate(&t0); cudaEventCreate(&t1); cudaEventCreate(&fork); cudaEventCreate(&ea); cudaEventCreate(&eb);
  cudaFuncSetAttribute(spin, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(t0, 0); cudaEventRecord(fork, 0);
    cudaStreamWaitEvent(a, fork, 0); cudaStreamWaitEvent(b, fork, 0);
    if (gridA) spin<<<gridA, 32, smem, a>>>(1000000, nullptr);
    if (ev_between) cudaEventRecord(ea, a);
    if (gridB) spin<<<gridB, 32, smem, b>>>(1000000, nullptr);
    if (!ev_between) cudaEventRecord(ea, a);
    cudaEventRecord(eb, b);
    cudaStreamWaitEvent(0, ea, 0); cudaStreamWaitEvent(0, eb, 0);
    cudaEventRecord(t1, 0); cudaEventSynchronize(t1);
    cudaEventElapsedTime(&ms, t0, t1);
  }
  return ms;
}
// The launch pattern of schedule mode 3 with spin kernels: A = 1628 pieces of j/12 units (j = 1 + i/148), B = 3492 whole units,
// C = the 1628 remainders (12-j)/12 after A.  5120 units on 1776 slots: 2.88 units if later grids backfill freed slots, 3.84 if
// the three grids run one after the other.
__global__ void spin_piece(long long unit, int per, int remainder) {
  const int j = 1 + blockIdx.x / per;
  const long long cycles = unit * (remainder ? 12 - j : j) / 12;
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {}
}
static float pattern3(bool prio, int smem, long long unit) {
  int least = 0, greatest = 0;
  cudaDeviceGetStreamPriorityRange(&least, &greatest);
  cudaStream_t a, b, c;
  cudaStreamCreateWithPriority(&a, cudaStreamNonBlocking, prio ? greatest : least);
  cudaStreamCreateWithPriority(&b, cudaStreamNonBlocking, prio ? greatest + 1 : least);
  cudaStreamCreateWithPriority(&c, cudaStreamNonBlocking, least);
  cudaEvent_t t0, t1, fork, ea, eb, ec;
  cudaEventCreate(&t0); cudaEventCreate(&t1); cudaEventCreate(&fork); cudaEventCreate(&ea); cudaEventCreate(&eb); cudaEventCreate(&ec);
  cudaFuncSetAttribute(spin, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(spin_piece, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(t0, 0); cudaEventRecord(fork, 0);
    cudaStreamWaitEvent(a, fork, 0); cudaStreamWaitEvent(b, fork, 0);
    spin_piece<<<1628, 32, smem, a>>>(unit, 148, 0);
    spin<<<3492, 32, smem, b>>>(unit, nullptr);
    cudaEventRecord(ea, a);
    cudaStreamWaitEvent(c, ea, 0);
    spin_piece<<<1628, 32, smem, c>>>(unit, 148, 1);
    cudaEventRecord(eb, b); cudaEventRecord(ec, c);
    cudaStreamWaitEvent(0, eb, 0); cudaStreamWaitEvent(0, ec, 0);
    cudaEventRecord(t1, 0); cudaEventSynchronize(t1);
    cudaEventElapsedTime(&ms, t0, t1);
  }
  return ms;
}

int main() {
  const char *mc = getenv("CUDA_DEVICE_MAX_CONNECTIONS");
  printf("CUDA_DEVICE_MAX_CONNECTIONS=%s\n", mc ? mc : "(unset)");
  for (int smem : {0, 17696}) {
    printf("smem %d: A alone (148 CTAs) %.3f ms, B alone %.3f ms, A+B event-between %.3f ms, A+B launches first %.3f ms; full machine A (1628) + B (1776): %.3f / %.3f ms\n", smem,
           run(148, 0, false, smem), run(0, 148, false, smem), run(148, 148, true, smem), run(148, 148, false, smem), run(1628, 1776, true, smem), run(1628, 1776, false, smem));
  }
  const float u = run(148, 0, false, 17696);
  printf("mode-3 pattern, unit = %.3f ms: priorities on %.3f ms, off %.3f ms (2.88 units = %.3f ms, 3.84 units = %.3f ms)\n", u,
         pattern3(true, 17696, 1000000), pattern3(false, 17696, 1000000), 2.88f * u, 3.84f * u);
  return 0;
}
