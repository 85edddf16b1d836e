/*
 * gpu_ref_harness.cu -- TEST / MEASUREMENT INFRASTRUCTURE (oracle side), not product code.
 *
 * Launches the reference's OWN CUDA kernels on the GPU box so that the B200-native kernel has a same-device
 * number to beat.  oracle/Makefile compiles this file together with ONE
 * of the UNMODIFIED reference sources, where it lies under /root/reference:
 *
 *   src/cloudsc_cuda/cloudsc/cloudsc_c_k_caching.cu  -> oracle/_ref/libcloudsc_gpuref_kcaching.so   ("k-caching")
 *   src/cloudsc_cuda/cloudsc/cloudsc_c.cu            -> oracle/_ref/libcloudsc_gpuref_scc.so        ("SCC")
 *
 * Both define `__global__ void cloudsc_c(...)` with the prototype of src/cloudsc_cuda/cloudsc/cloudsc_c.h, which is
 * included here from the reference tree (-I).  The launch is the reference driver's
 * (src/cloudsc_cuda/cloudsc/cloudsc_driver.cu:391-416): block = (NPROMA,1,1), grid = (1,1,nblocks), kidia = 1,
 * kfdia = min(NPROMA, NGPTOT), the parameter struct in device global memory (:383), the 28 constants by value.
 * Timed with CUDA events around the launch only; PLUDE (in/out, cloudsc_c.c:970) is restored from a pristine copy
 * before every launch, outside the event pair -- the same protocol bench.py applies to the product kernel.
 *
 * Never linked into, loaded by or reachable from dwarf_p_cloudsc_b200/.
 */
#include <cstdio>
#include <cuda_runtime.h>

#include "cloudsc_c.h" /* the reference's header: kernel prototype + struct TECLDP (yoecldp_c.h) */

#define NF 56 /* field table in cloudsc_c() argument order of the C dwarf (include/cloudsc_b200.h) */
enum {
  F_PT, F_PQ, F_CML_T, F_CML_Q, F_CML_A, F_CML_CLD, F_TMP_T, F_TMP_Q, F_TMP_A, F_TMP_CLD, F_LOC_T, F_LOC_Q, F_LOC_A,
  F_LOC_CLD, F_PVFA, F_PVFL, F_PVFI, F_PDYNA, F_PDYNL, F_PDYNI, F_PHRSW, F_PHRLW, F_PVERVEL, F_PAP, F_PAPH, F_PLSM,
  F_KTYPE, F_PLU, F_PLUDE, F_PSNDE, F_PMFU, F_PMFD, F_PA, F_PCLV, F_PSUPSAT, F_PLCRIT_AER, F_PICRIT_AER, F_PRE_ICE,
  F_PCCN, F_PNICE, F_PCOVPTOT, F_PRAINFRAC, F_PFSQLF, F_PFSQIF, F_PFCQNNG, F_PFCQLNG, F_PFSQRF, F_PFSQSF, F_PFCQRNG,
  F_PFCQSNG, F_PFSQLTUR, F_PFSQITUR, F_PFPLSL, F_PFPLSN, F_PFHPSL, F_PFHPSN
};

extern "C" {

/* sizeof(struct TECLDP) as the reference's CUDA tree declares it (must equal the C dwarf's 912 bytes) */
int cloudsc_gpuref_tecldp_bytes(void) { return (int)sizeof(struct TECLDP); }

/*
 * dev[]      56 device pointers, FP64 fields in the driver's [nblocks][..][nproma] layout (never-touched fields may be NULL)
 * tecldp     host copy of the parameter struct (uploaded to device global memory, as the reference driver does)
 * k[28]      rg rd rcpd retv rlvtt rlstt rlmlt rtt rv r2es r3les r3ies r4les r4ies r5les r5ies r5alvcp r5alscp ralvdcp
 *            ralsdcp ralfdcp rtwat rtice rticecu rtwat_rtice_r rtwat_rticecu_r rkoop1 rkoop2
 * plude0     device copy of the pristine PLUDE (restored before every launch), may be NULL when warmup + reps == 1
 * ms[3]      out: mean, min, max kernel time over `reps` timed launches
 * returns 0 or a cudaError_t
 */
int cloudsc_gpuref_run(void *const dev[NF], const void *tecldp, const double k[28], int ngptot, int nproma, double ptsphy,
                       const void *plude0, int warmup, int reps, float ms[3]) {
  if (ngptot <= 0 || nproma <= 0 || nproma > 1024 || reps < 1) return (int)cudaErrorInvalidValue;
  const int nblocks = (ngptot + nproma - 1) / nproma;
  const int klev = 137; /* hard-coded in the reference kernels (cloudsc_c_k_caching.cu:53) */
  struct TECLDP *d_y = nullptr;
  cudaError_t e = cudaMalloc(&d_y, sizeof(struct TECLDP));
  if (e != cudaSuccess) return (int)e;
  cudaMemcpy(d_y, tecldp, sizeof(struct TECLDP), cudaMemcpyHostToDevice);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const dim3 blockdim(nproma, 1, 1), griddim(1, 1, nblocks);
  const int icend = nproma < ngptot ? nproma : ngptot;
  const size_t plude_bytes = (size_t)nblocks * klev * nproma * sizeof(double);
#define D(i) ((double *)dev[i])
  double sum = 0.0;
  float mn = 1e30f, mx = 0.f;
  for (int it = 0; it < warmup + reps; ++it) {
    if (plude0 && it > 0) cudaMemcpyAsync(dev[F_PLUDE], plude0, plude_bytes, cudaMemcpyDeviceToDevice, 0);
    cudaEventRecord(a, 0);
    cloudsc_c<<<griddim, blockdim>>>(1, icend, nproma, ptsphy, D(F_PT), D(F_PQ), D(F_TMP_T), D(F_TMP_Q), D(F_TMP_A), D(F_TMP_CLD),
        D(F_LOC_T), D(F_LOC_Q), D(F_LOC_A), D(F_LOC_CLD), D(F_PVFA), D(F_PVFL), D(F_PVFI), D(F_PDYNA), D(F_PDYNL), D(F_PDYNI),
        D(F_PHRSW), D(F_PHRLW), D(F_PVERVEL), D(F_PAP), D(F_PAPH), D(F_PLSM), (const int *)dev[F_KTYPE], D(F_PLU), D(F_PLUDE),
        D(F_PSNDE), D(F_PMFU), D(F_PMFD), D(F_PA), D(F_PCLV), D(F_PSUPSAT), D(F_PLCRIT_AER), D(F_PICRIT_AER), D(F_PRE_ICE),
        D(F_PCCN), D(F_PNICE), D(F_PCOVPTOT), D(F_PRAINFRAC), D(F_PFSQLF), D(F_PFSQIF), D(F_PFCQNNG), D(F_PFCQLNG), D(F_PFSQRF),
        D(F_PFSQSF), D(F_PFCQRNG), D(F_PFCQSNG), D(F_PFSQLTUR), D(F_PFSQITUR), D(F_PFPLSL), D(F_PFPLSN), D(F_PFHPSL),
        D(F_PFHPSN), d_y, nblocks, k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7], k[8], k[9], k[10], k[11], k[12], k[13], k[14],
        k[15], k[16], k[17], k[18], k[19], k[20], k[21], k[22], k[23], k[24], k[25], k[26], k[27]);
    cudaEventRecord(b, 0);
    e = cudaEventSynchronize(b);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) break;
    float t = 0.f;
    cudaEventElapsedTime(&t, a, b);
    if (it >= warmup) { sum += t; mn = t < mn ? t : mn; mx = t > mx ? t : mx; }
  }
#undef D
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d_y);
  if (e != cudaSuccess) return (int)e;
  if (ms) { ms[0] = (float)(sum / reps); ms[1] = mn; ms[2] = mx; }
  return 0;
}

/* registers / local bytes / max threads of the reference kernel as compiled for this device */
int cloudsc_gpuref_kernel_info(int out[4]) {
  cudaFuncAttributes a;
  cudaError_t e = cudaFuncGetAttributes(&a, (const void *)cloudsc_c);
  if (e != cudaSuccess) return (int)e;
  out[0] = a.numRegs; out[1] = (int)a.localSizeBytes; out[2] = (int)a.sharedSizeBytes; out[3] = a.maxThreadsPerBlock;
  return 0;
}
}
