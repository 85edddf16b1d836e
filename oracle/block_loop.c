/*
 * block_loop.c -- TEST INFRASTRUCTURE (oracle side), not product code.
 *
 * The OpenMP loop over NPROMA blocks of the reference driver
 * (src/cloudsc_c/cloudsc/cloudsc_driver.c:183-229) restated over a 56-entry field table, so that
 * a CPU kernel with the cloudsc_c() signature can be run over all blocks from Python/ctypes:
 * block offsets as :193-196, bsize as :197, pcovptot / tendency_loc_cld zeroed as :199-200.
 *
 * Compiled twice by oracle/Makefile:
 *   -DORACLE_KERNEL=cloudsc_c      -DORACLE_BLOCKS=cloudsc_ref_blocks     with the UNMODIFIED reference
 *                                   kernel sources  -> oracle/_ref/libcloudsc_ref.so
 *   -DORACLE_KERNEL=cloudsc_oracle -DORACLE_BLOCKS=cloudsc_oracle_blocks  with oracle/cloudsc_oracle.c
 *                                                   -> oracle/libcloudsc_oracle.so
 * Returns the wall time of the loop in seconds (omp_get_wtime), <0 on bad arguments.
 */
#include <stddef.h>
#include <omp.h>
#include "cloudsc_b200.h"

int ORACLE_KERNEL(int kidia, int kfdia, int klon, int klev, double ptsphy,
  double *, double *, double *, double *, double *, double *, double *, double *, double *, double *,
  double *, double *, double *, double *, double *, double *, double *, double *, double *, double *,
  double *, double *, double *, double *, double *, double *, int *, double *, double *, double *,
  double *, double *, double *, double *, double *, double *, double *, double *, double *, double *,
  double *, double *, double *, double *, double *, double *, double *, double *, double *, double *,
  double *, double *, double *, double *, double *, double *);

static const char kinds[CLOUDSC_NFIELDS + 1] = {
#define CLOUDSC_X(name, kind, use) #kind[0],
  CLOUDSC_FIELDS(CLOUDSC_X)
#undef CLOUDSC_X
  0};

static size_t block_elems(int f, int nproma, int klev) {
  switch (kinds[f]) {
    case 'L': return (size_t)klev * nproma;
    case 'H': return (size_t)(klev + 1) * nproma;
    case 'C': return (size_t)5 * klev * nproma;
    default:  return (size_t)nproma;
  }
}

double ORACLE_BLOCKS(int nthreads, int ngptot, int nproma, int klev, double ptsphy,
                     void *const fields[CLOUDSC_NFIELDS]) {
  if (ngptot <= 0 || nproma <= 0 || klev <= 0) return -1.0;
  const int nblocks = ngptot / nproma + (ngptot % nproma ? 1 : 0);
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  const double t0 = omp_get_wtime();
#pragma omp parallel for schedule(runtime) num_threads(nthreads)
  for (int b = 0; b < nblocks; b++) {
    void *p[CLOUDSC_NFIELDS];
    for (int f = 0; f < CLOUDSC_NFIELDS; f++) {
      const size_t off = (size_t)b * block_elems(f, nproma, klev);
      p[f] = fields[f] ? (void *)((char *)fields[f] + off * (kinds[f] == 'I' ? sizeof(int) : sizeof(double))) : NULL;
    }
    const int rest = ngptot - b * nproma;
    const int bsize = rest < nproma ? rest : nproma;
    double *cov = (double *)p[CLOUDSC_F_PCOVPTOT], *tcld = (double *)p[CLOUDSC_F_TENDENCY_LOC_CLD];
    for (size_t i = 0; i < (size_t)klev * nproma; i++) cov[i] = 0.0;
    for (size_t i = 0; i < (size_t)5 * klev * nproma; i++) tcld[i] = 0.0;
#define D(i) ((double *)p[i])
    ORACLE_KERNEL(1, bsize, nproma, klev, ptsphy,
      D(0), D(1), D(2), D(3), D(4), D(5), D(6), D(7), D(8), D(9), D(10), D(11), D(12), D(13), D(14), D(15),
      D(16), D(17), D(18), D(19), D(20), D(21), D(22), D(23), D(24), D(25), (int *)p[26], D(27), D(28), D(29),
      D(30), D(31), D(32), D(33), D(34), D(35), D(36), D(37), D(38), D(39), D(40), D(41), D(42), D(43), D(44),
      D(45), D(46), D(47), D(48), D(49), D(50), D(51), D(52), D(53), D(54), D(55));
#undef D
  }
  return omp_get_wtime() - t0;
}
