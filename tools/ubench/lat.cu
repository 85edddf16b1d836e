// Dependent-issue latency / throughput microbenchmarks for the FP64 instruction mix of CLOUDSC on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu ; run: ./lat
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int MODE> __global__ void chain(double *out, long long *cyc, double a, double b, int ilp) {
  double x0 = a + threadIdx.x, x1 = a * 2, x2 = a * 3, x3 = a * 4;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N / 16; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (MODE == 0) { x0 = fma(x0, b, a); }                                   // 1 dependent DFMA chain
      if (MODE == 1) { x0 = fma(x0, b, a); x1 = fma(x1, b, a); }               // 2 chains
      if (MODE == 2) { x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a); }
      if (MODE == 3) { x0 = x0 + b; }                                          // DADD
      if (MODE == 4) { x0 = x0 * b; }                                          // DMUL
      if (MODE == 5) { x0 = x0 > b ? x0 : b + x0; }                            // DSETP + select + DADD
      if (MODE == 6) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x0)); x0 = r + b; } // MUFU.RCP64H + DADD
      if (MODE == 7) { x0 = sqrt(x0) + b; }
      if (MODE == 8) { x0 = 1.0 / x0 + b; }
      if (MODE == 9) { x0 = exp(x0 * 1e-3) ; }
      if (MODE == 10) { x0 = pow(x0 + 1.5, 0.666) ; }
      if (MODE == 11) { x0 = log(x0 + 1.5) ; }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char *name, int ops_per_iter, int threads) {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 8 * 1024 * 8); cudaMalloc(&cyc, 8 * 64);
  chain<MODE><<<1, threads>>>(out, cyc, 1.0000001, 0.9999999, 1);
  chain<MODE><<<1, threads>>>(out, cyc, 1.0000001, 0.9999999, 1);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s threads %4d: %7.2f cycles per step (%d op/step)\n", name, threads, (double)h / N, ops_per_iter);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int th : {32, 128, 256, 512}) {
    run<0>("DFMA chain x1", 1, th); run<1>("DFMA chain x2", 2, th); run<2>("DFMA chain x4", 4, th);
  }
  run<3>("DADD chain", 1, 32); run<4>("DMUL chain", 1, 32); run<5>("DSETP+SEL+DADD", 1, 32);
  run<6>("MUFU.RCP64H + DADD", 1, 32); run<7>("sqrt + DADD", 1, 32); run<8>("1/x + DADD", 1, 32);
  run<9>("exp()", 1, 32); run<10>("pow(x,.666)", 1, 32); run<11>("log()", 1, 32);
  run<9>("exp() 128thr", 1, 128); run<10>("pow(x,.666) 128thr", 1, 128);
  return 0;
}
