// Host<->device copy bandwidth, 1..N GPUs AT THE SAME TIME: H2D alone, D2H alone, both directions at once, for three
// kinds of page-locked host memory.  The end-to-end path of the library (cloudsc_b200_run_host) is bound by these
// numbers; bench.py's e2e figure is reported against them (profiles/r2_pcie_concurrency.txt).
//
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Xcompiler -pthread -o pcie pcie.cu
//   run:   ./pcie [--gpus N] [--mb 1024] [--reps 3] [--mode hostalloc|wc|thp|all] [--chunk-mb 0] [--bind 0|1]
//
//   hostalloc  cudaHostAlloc(cudaHostAllocPortable)                      (what cloudsc_b200_host_alloc / torch pin_memory give)
//   wc         the H2D source additionally cudaHostAllocWriteCombined    (uncached on the CPU side: no snoops on device reads)
//   thp        aligned_alloc(2 MB) + madvise(MADV_HUGEPAGE) + first touch by the GPU's own host thread + cudaHostRegister
//              (2 MB pages: 512x fewer IOMMU translations; first touch places the pages on the toucher's NUMA node)
//   --chunk-mb k   issue every transfer as k-MB pieces (what a chunked pipeline does) instead of one copy
//   --bind 1       pin host thread g to the g-th slice of the online CPUs before it allocates/touches its buffers
// One host thread per GPU; every phase starts behind a barrier and is timed per GPU with CUDA events; the aggregate is
// total bytes / slowest GPU.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>
#include <cuda_runtime.h>

static pthread_barrier_t g_bar;
struct Result { float h2d = 0, d2h = 0, bi_h2d = 0, bi_d2h = 0, bi_wall = 0; int ok = 0; };

static void *alloc_host(const std::string &mode, size_t n, bool upload_side) {
  void *p = nullptr;
  if (mode == "thp") {
    p = aligned_alloc(2u << 20, n);
    if (!p) return nullptr;
    madvise(p, n, MADV_HUGEPAGE);
    memset(p, 1, n);   // first touch from this thread
    if (cudaHostRegister(p, n, cudaHostRegisterPortable) != cudaSuccess) return nullptr;
    return p;
  }
  unsigned flags = cudaHostAllocPortable;
  if (mode == "wc" && upload_side) flags |= cudaHostAllocWriteCombined;
  if (cudaHostAlloc(&p, n, flags) != cudaSuccess) return nullptr;
  memset(p, 1, n);
  return p;
}

static void copy(char *dst, const char *src, size_t n, size_t chunk, cudaMemcpyKind kind, cudaStream_t s) {
  if (!chunk || chunk >= n) { cudaMemcpyAsync(dst, src, n, kind, s); return; }
  for (size_t o = 0; o < n; o += chunk) cudaMemcpyAsync(dst + o, src + o, std::min(chunk, n - o), kind, s);
}

static void worker(int g, int ngpu, const std::string &mode, size_t n, size_t chunk, int reps, bool bind, Result *out) {
  if (bind) {
    const int ncpu = (int)sysconf(_SC_NPROCESSORS_ONLN), per = std::max(1, ncpu / ngpu);
    cpu_set_t set; CPU_ZERO(&set);
    for (int c = g * per; c < std::min(ncpu, (g + 1) * per); ++c) CPU_SET(c, &set);
    sched_setaffinity(0, sizeof(set), &set);
  }
  cudaSetDevice(g);
  char *h_up = (char *)alloc_host(mode, n, true), *h_dn = (char *)alloc_host(mode, n, false), *d_up = nullptr, *d_dn = nullptr;
  cudaMalloc(&d_up, n); cudaMalloc(&d_dn, n);
  cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  cudaEvent_t a1, b1, a2, b2; cudaEventCreate(&a1); cudaEventCreate(&b1); cudaEventCreate(&a2); cudaEventCreate(&b2);
  const bool ok = h_up && h_dn && d_up && d_dn;
  if (ok) { cudaMemcpy(d_up, h_up, n, cudaMemcpyHostToDevice); cudaMemcpy(h_dn, d_dn, n, cudaMemcpyDeviceToHost); }  // warm-up
  auto gbs = [&](float ms) { return ms > 0 ? (float)((double)n * reps / (ms * 1e6)) : 0.f; };
  float ms;
  // --- H2D alone
  pthread_barrier_wait(&g_bar);
  if (ok) {
    cudaEventRecord(a1, s1);
    for (int r = 0; r < reps; ++r) copy(d_up, h_up, n, chunk, cudaMemcpyHostToDevice, s1);
    cudaEventRecord(b1, s1); cudaEventSynchronize(b1); cudaEventElapsedTime(&ms, a1, b1); out->h2d = gbs(ms);
  }
  // --- D2H alone
  pthread_barrier_wait(&g_bar);
  if (ok) {
    cudaEventRecord(a2, s2);
    for (int r = 0; r < reps; ++r) copy(h_dn, d_dn, n, chunk, cudaMemcpyDeviceToHost, s2);
    cudaEventRecord(b2, s2); cudaEventSynchronize(b2); cudaEventElapsedTime(&ms, a2, b2); out->d2h = gbs(ms);
  }
  // --- both directions
  pthread_barrier_wait(&g_bar);
  if (ok) {
    cudaEventRecord(a1, s1); cudaEventRecord(a2, s2);
    for (int r = 0; r < reps; ++r) {
      copy(d_up, h_up, n, chunk, cudaMemcpyHostToDevice, s1);
      copy(h_dn, d_dn, n, chunk, cudaMemcpyDeviceToHost, s2);
    }
    cudaEventRecord(b1, s1); cudaEventRecord(b2, s2);
    cudaEventSynchronize(b1); cudaEventSynchronize(b2);
    cudaEventElapsedTime(&ms, a1, b1); out->bi_h2d = gbs(ms);
    float ms2; cudaEventElapsedTime(&ms2, a2, b2); out->bi_d2h = gbs(ms2);
    out->bi_wall = std::max(ms, ms2);
  }
  pthread_barrier_wait(&g_bar);
  out->ok = ok;
  if (mode == "thp") { if (h_up) { cudaHostUnregister(h_up); free(h_up); } if (h_dn) { cudaHostUnregister(h_dn); free(h_dn); } }
  else { if (h_up) cudaFreeHost(h_up); if (h_dn) cudaFreeHost(h_dn); }
  cudaFree(d_up); cudaFree(d_dn);
}

// --pattern 1: the access pattern of cloudsc_b200_run_host on ONE GPU: `nin` separate upload arrays and `nout` separate
// download arrays of `mb` MB each, walked chunk by chunk (one copy of chunk_mb MB per array per chunk; uploads on one
// stream, downloads on another, no dependencies), against the same bytes moved as two flat buffers.
static void pattern(const std::string &mode, size_t arr_bytes, size_t chunk, int nin, int nout) {
  cudaSetDevice(0);
  std::vector<char *> hin(nin), hout(nout);
  for (auto &p : hin) p = (char *)alloc_host(mode, arr_bytes, true);
  for (auto &p : hout) p = (char *)alloc_host(mode, arr_bytes, false);
  char *d_up, *d_dn;
  const size_t ring = 4 * chunk * (size_t)std::max(nin, nout);
  cudaMalloc(&d_up, ring); cudaMalloc(&d_dn, ring);
  cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  cudaEvent_t a1, b1, a2, b2; cudaEventCreate(&a1); cudaEventCreate(&b1); cudaEventCreate(&a2); cudaEventCreate(&b2);
  const int nchunks = (int)(arr_bytes / chunk);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a1, s1); cudaEventRecord(a2, s2);
    for (int c = 0; c < nchunks; ++c) {
      char *du = d_up + (size_t)(c % 4) * chunk * nin, *dd = d_dn + (size_t)(c % 4) * chunk * nout;
      for (int f = 0; f < nin; ++f) cudaMemcpyAsync(du + (size_t)f * chunk, hin[f] + (size_t)c * chunk, chunk, cudaMemcpyHostToDevice, s1);
      for (int f = 0; f < nout; ++f) cudaMemcpyAsync(hout[f] + (size_t)c * chunk, dd + (size_t)f * chunk, chunk, cudaMemcpyDeviceToHost, s2);
    }
    cudaEventRecord(b1, s1); cudaEventRecord(b2, s2);
    cudaEventSynchronize(b1); cudaEventSynchronize(b2);
    float m1, m2; cudaEventElapsedTime(&m1, a1, b1); cudaEventElapsedTime(&m2, a2, b2);
    const double up = (double)nin * nchunks * chunk, dn = (double)nout * nchunks * chunk;
    printf("  pattern %-9s rep %d: %d+%d arrays x %zu MB in %zu-KB pieces: H2D %.2f GB in %.1f ms (%.1f GB/s), D2H %.2f GB in %.1f ms (%.1f GB/s), sum %.1f GB/s over %.1f ms\n",
           mode.c_str(), rep, nin, nout, arr_bytes >> 20, chunk >> 10, up / 1e9, m1, up / m1 / 1e6, dn / 1e9, m2, dn / m2 / 1e6,
           (up + dn) / std::max(m1, m2) / 1e6, std::max(m1, m2));
  }
}


// --pattern 2 --knobs K: pattern 1 made to look like cloudsc_b200_run_host step by step, to find which property of the real
// pipeline costs bandwidth.  K is a bit mask:
//   1  ring dependencies: upload c waits for download c-4 (its staging buffer), download c waits for upload c
//   2  piece = 54 NPROMA blocks of a [137][128] double field = 7 575 552 B (a multiple of 1 KB, not of 4 KB) instead of 8 MB
//   4  every copy is a cudaMemcpy2DAsync of 54 rows: pitch = one block (140 288 B), width = 124/137 of it (row trimming)
//   8  a host thread memsets 180 MB of pageable memory while the copies run (species-5 zeroing)
//  16  an empty kernel on a third stream between upload c and download c
__global__ void nop_kernel() {}
static void pattern2(const std::string &mode, int knobs, int nin, int nout, int nchunks) {
  cudaSetDevice(0);
  const size_t blk = 140288, nb = 54;
  const size_t piece = (knobs & 2) || (knobs & 4) ? nb * blk : (size_t)8 << 20;
  const size_t width = (knobs & 4) ? (size_t)124 * 1024 : blk;
  const size_t arr_bytes = piece * nchunks;
  std::vector<char *> hin(nin), hout(nout);
  for (auto &p : hin) p = (char *)alloc_host(mode, arr_bytes, true);
  for (auto &p : hout) p = (char *)alloc_host(mode, arr_bytes, false);
  char *d_up, *d_dn;
  cudaMalloc(&d_up, 4 * piece * nin); cudaMalloc(&d_dn, 4 * piece * nout);
  cudaStream_t s1, s2, s3;
  cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s3, cudaStreamNonBlocking);
  cudaEvent_t a1, b1, a2, b2, up[4], dn[4], kk[4];
  cudaEventCreate(&a1); cudaEventCreate(&b1); cudaEventCreate(&a2); cudaEventCreate(&b2);
  for (int i = 0; i < 4; ++i) { cudaEventCreateWithFlags(&up[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&dn[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&kk[i], cudaEventDisableTiming); }
  char *scratch = (char *)malloc((size_t)180 << 20);
  for (int rep = 0; rep < 3; ++rep) {
    std::thread ms;
    cudaEventRecord(a1, s1); cudaEventRecord(a2, s2);
    for (int c = 0; c < nchunks; ++c) {
      const int l = c % 4;
      char *du = d_up + (size_t)l * piece * nin, *dd = d_dn + (size_t)l * piece * nout;
      if ((knobs & 1) && c >= 4) cudaStreamWaitEvent(s1, dn[l], 0);
      for (int f = 0; f < nin; ++f) {
        if (knobs & 4) cudaMemcpy2DAsync(du + (size_t)f * piece, blk, hin[f] + (size_t)c * piece, blk, width, nb, cudaMemcpyHostToDevice, s1);
        else cudaMemcpyAsync(du + (size_t)f * piece, hin[f] + (size_t)c * piece, piece, cudaMemcpyHostToDevice, s1);
      }
      if (knobs & 1) cudaEventRecord(up[l], s1);
      if (knobs & 16) {
        if (knobs & 1) cudaStreamWaitEvent(s3, up[l], 0);
        nop_kernel<<<1, 32, 0, s3>>>();
        cudaEventRecord(kk[l], s3);
        if (knobs & 1) cudaStreamWaitEvent(s2, kk[l], 0);
      } else if (knobs & 1) cudaStreamWaitEvent(s2, up[l], 0);
      for (int f = 0; f < nout; ++f) {
        if (knobs & 4) cudaMemcpy2DAsync(hout[f] + (size_t)c * piece, blk, dd + (size_t)f * piece, blk, width, nb, cudaMemcpyDeviceToHost, s2);
        else cudaMemcpyAsync(hout[f] + (size_t)c * piece, dd + (size_t)f * piece, piece, cudaMemcpyDeviceToHost, s2);
      }
      if (knobs & 1) cudaEventRecord(dn[l], s2);
    }
    if (knobs & 8) ms = std::thread([&] { memset(scratch, rep, (size_t)180 << 20); });
    cudaEventRecord(b1, s1); cudaEventRecord(b2, s2);
    cudaEventSynchronize(b1); cudaEventSynchronize(b2);
    if (ms.joinable()) ms.join();
    float m1, m2; cudaEventElapsedTime(&m1, a1, b1); cudaEventElapsedTime(&m2, a2, b2);
    const double per = (double)((knobs & 4) ? width * nb : piece);
    const double upb = (double)nin * nchunks * per, dnb = (double)nout * nchunks * per;
    printf("  pattern2 %-9s knobs %2d rep %d: %d+%d arrays, %d chunks: H2D %.2f GB in %.1f ms (%.1f GB/s), D2H %.2f GB in %.1f ms (%.1f GB/s), sum %.1f GB/s over %.1f ms\n",
           mode.c_str(), knobs, rep, nin, nout, nchunks, upb / 1e9, m1, upb / m1 / 1e6, dnb / 1e9, m2, dnb / m2 / 1e6, (upb + dnb) / std::max(m1, m2) / 1e6, std::max(m1, m2));
  }
  for (auto p : hin) cudaFreeHost(p);
  for (auto p : hout) cudaFreeHost(p);
  cudaFree(d_up); cudaFree(d_dn); free(scratch);
}

int main(int argc, char **argv) {
  int ngpu = 1, reps = 3, bind = 0, pat = 0, nin = 35, nout = 22, knobs = 0, nchunks = 24;
  size_t mb = 1024, chunk_mb = 0;
  std::string mode = "all";
  for (int i = 1; i + 1 < argc; i += 2) {
    std::string a = argv[i];
    if (a == "--gpus") ngpu = atoi(argv[i + 1]);
    else if (a == "--mb") mb = (size_t)atol(argv[i + 1]);
    else if (a == "--reps") reps = atoi(argv[i + 1]);
    else if (a == "--mode") mode = argv[i + 1];
    else if (a == "--chunk-mb") chunk_mb = (size_t)atol(argv[i + 1]);
    else if (a == "--bind") bind = atoi(argv[i + 1]);
    else if (a == "--pattern") pat = atoi(argv[i + 1]);
    else if (a == "--nin") nin = atoi(argv[i + 1]);
    else if (a == "--nout") nout = atoi(argv[i + 1]);
    else if (a == "--knobs") knobs = atoi(argv[i + 1]);
    else if (a == "--nchunks") nchunks = atoi(argv[i + 1]);
  }
  int have = 0;
  cudaGetDeviceCount(&have);
  if (have < 1) { fprintf(stderr, "no CUDA device\n"); return 1; }
  if (ngpu > have) ngpu = have;
  if (pat == 2) { pattern2(mode == "all" ? "hostalloc" : mode, knobs, nin, nout, nchunks); return 0; }
  if (pat) {
    for (const std::string &m : (mode == "all" ? std::vector<std::string>{"hostalloc", "thp"} : std::vector<std::string>{mode}))
      pattern(m, mb << 20, (chunk_mb ? chunk_mb : 8) << 20, nin, nout);
    return 0;
  }
  std::vector<std::string> modes = mode == "all" ? std::vector<std::string>{"hostalloc", "wc", "thp"} : std::vector<std::string>{mode};
  printf("# %d GPU(s) concurrently, %zu MB per direction per GPU x %d reps, chunk %zu MB, bind %d\n", ngpu, mb, reps, chunk_mb, bind);
  printf("# %-10s %4s %10s %10s %12s %12s %12s\n", "mode", "gpu", "H2D GB/s", "D2H GB/s", "bi H2D GB/s", "bi D2H GB/s", "bi sum GB/s");
  for (const std::string &m : modes) {
    std::vector<Result> res(ngpu);
    pthread_barrier_init(&g_bar, nullptr, ngpu);
    std::vector<std::thread> th;
    for (int g = 0; g < ngpu; ++g) th.emplace_back(worker, g, ngpu, m, mb << 20, chunk_mb << 20, reps, bind != 0, &res[g]);
    for (auto &t : th) t.join();
    pthread_barrier_destroy(&g_bar);
    double s_h2d = 0, s_d2h = 0, s_bi = 0;
    float slow_h2d = 1e30f, slow_d2h = 1e30f, slow_bi = 0.f;
    for (int g = 0; g < ngpu; ++g) {
      const Result &r = res[g];
      if (!r.ok) { printf("  %-10s %4d   allocation failed\n", m.c_str(), g); continue; }
      printf("  %-10s %4d %10.1f %10.1f %12.1f %12.1f %12.1f\n", m.c_str(), g, r.h2d, r.d2h, r.bi_h2d, r.bi_d2h, r.bi_h2d + r.bi_d2h);
      s_h2d += r.h2d; s_d2h += r.d2h; s_bi += r.bi_h2d + r.bi_d2h;
      slow_h2d = std::min(slow_h2d, r.h2d); slow_d2h = std::min(slow_d2h, r.d2h); slow_bi = std::max(slow_bi, r.bi_wall);
    }
    const double bytes = (double)(mb << 20) * reps;
    printf("  %-10s  sum %10.1f %10.1f %38.1f   | slowest-GPU aggregate: H2D %.1f  D2H %.1f  both %.1f GB/s\n", m.c_str(), s_h2d, s_d2h, s_bi,
           ngpu * slow_h2d, ngpu * slow_d2h, slow_bi > 0 ? 2.0 * ngpu * bytes / (slow_bi * 1e6) : 0.0);
  }
  return 0;
}
