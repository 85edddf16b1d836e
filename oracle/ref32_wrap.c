/*
 * ref32_wrap.c -- TEST INFRASTRUCTURE (oracle side), not product code.
 *
 * A SINGLE-PRECISION build of the reference's own kernel, as the oracle of the FP32 device path (BASELINE configs[4]).
 * The reference C sources are hard-wired `double` (only the Fortran family has a --single-precision build, JPRB=real32,
 * src/common/module/parkind1.F90:40-44, and its CI never runs it, .github/workflows/build.yml:172), so this translation
 * unit re-types the UNMODIFIED files where they lie under /root/reference by the preprocessor, the way JPRB=real32 does
 * for the Fortran kernel:
 *     double -> float, exp/pow/sqrt/fabs/fmax/fmin/copysign -> their float twins, DBL_EPSILON -> FLT_EPSILON,
 *     every floating literal single (gcc -fsingle-precision-constant, oracle/Makefile),
 * after the system headers have been seen with their real types.  Included verbatim: cloudsc/cloudsc_c.c,
 * yoecldp_c.c, yomcst_c.c, yoethf_c.c.  Plus the block loop of cloudsc_driver.c:183-229 over float fields and a setter
 * that narrows the double-precision parameter set into the re-typed globals.
 *   -> oracle/_ref/libcloudsc_ref32.so   (make -C oracle ref32)
 */
#include <float.h>
#include <math.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <omp.h>

typedef double ref32_f64;   /* the one place that still needs a binary64: the interface below */

#define double float
#define exp expf
#define pow powf
#define sqrt sqrtf
#define fabs fabsf
#define fmax fmaxf
#define fmin fminf
#define copysign copysignf
#undef DBL_EPSILON
#define DBL_EPSILON FLT_EPSILON

#include "cloudsc_c.c"
#include "yoecldp_c.c"
#include "yomcst_c.c"
#include "yoethf_c.c"

/* parameters: the 102 leading reals, the 17 ints, nshapep/nshapeq/nbeta of struct TECLDP in declaration order
 * (yoecldp_c.h:21-141), then the 28 YOMCST/YOETHF constants in the order of CLOUDSC_KERNEL_CONSTS (include/cloudsc_params.h) */
static struct TECLDP g_y32;
void cloudsc_ref32_set_params(const ref32_f64 *reals, int nreals, const int *ints, int nints, ref32_f64 nshapep, ref32_f64 nshapeq, int nbeta,
                              const ref32_f64 *k) {
  float *r = (float *)&g_y32;                         /* the reals lead the struct, contiguously */
  for (int i = 0; i < nreals; i++) r[i] = (float)reals[i];
  int *ip = (int *)(r + nreals);
  for (int i = 0; i < nints; i++) ip[i] = ints[i];
  g_y32.nshapep = (float)nshapep; g_y32.nshapeq = (float)nshapeq; g_y32.nbeta = nbeta;
  yrecldp = &g_y32;
  nclv = 5; ncldql = 1; ncldqi = 2; ncldqr = 3; ncldqs = 4; ncldqv = 5;
  rg = (float)k[0]; rd = (float)k[1]; rcpd = (float)k[2]; retv = (float)k[3]; rlvtt = (float)k[4]; rlstt = (float)k[5];
  rlmlt = (float)k[6]; rtt = (float)k[7]; rv = (float)k[8]; r2es = (float)k[9]; r3les = (float)k[10]; r3ies = (float)k[11];
  r4les = (float)k[12]; r4ies = (float)k[13]; r5les = (float)k[14]; r5ies = (float)k[15]; r5alvcp = (float)k[16];
  r5alscp = (float)k[17]; ralvdcp = (float)k[18]; ralsdcp = (float)k[19]; ralfdcp = (float)k[20]; rtwat = (float)k[21];
  rtice = (float)k[22]; rticecu = (float)k[23]; rtwat_rtice_r = (float)k[24]; rtwat_rticecu_r = (float)k[25];
  rkoop1 = (float)k[26]; rkoop2 = (float)k[27];
}
int cloudsc_ref32_sizeof_tecldp(void) { return (int)sizeof(struct TECLDP); }

/* field table in cloudsc_c() argument order; kinds as include/cloudsc_b200.h: L [klev], H [klev+1], C [5][klev], S/I [1] */
static const char k32_kinds[57] = "LLLLLCLLLCLLLCLLLLLLLLLLHSILLLLLLCLLLLLLLSHHHHHHHHHHHHHH";
static size_t k32_elems(int f, int nproma, int klev) {
  switch (k32_kinds[f]) {
    case 'L': return (size_t)klev * nproma;
    case 'H': return (size_t)(klev + 1) * nproma;
    case 'C': return (size_t)5 * klev * nproma;
    default: return (size_t)nproma;
  }
}

ref32_f64 cloudsc_ref32_blocks(int nthreads, int ngptot, int nproma, int klev, ref32_f64 ptsphy, void *const fields[56]) {
  if (ngptot <= 0 || nproma <= 0 || klev <= 0) return -1.0;
  const int nblocks = ngptot / nproma + (ngptot % nproma ? 1 : 0);
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  const ref32_f64 t0 = omp_get_wtime();
#pragma omp parallel for schedule(runtime) num_threads(nthreads)
  for (int b = 0; b < nblocks; b++) {
    void *p[56];
    for (int f = 0; f < 56; f++)
      p[f] = fields[f] ? (void *)((char *)fields[f] + (size_t)b * k32_elems(f, nproma, klev) * 4) : NULL;   /* float and int: 4 bytes */
    const int rest = ngptot - b * nproma;
    const int bsize = rest < nproma ? rest : nproma;
    float *cov = (float *)p[40], *tcld = (float *)p[13];
    for (size_t i = 0; i < (size_t)klev * nproma; i++) cov[i] = 0.0f;          /* cloudsc_driver.c:199-200 */
    for (size_t i = 0; i < (size_t)5 * klev * nproma; i++) tcld[i] = 0.0f;
#define D(i) ((float *)p[i])
    cloudsc_c(1, bsize, nproma, klev, (float)ptsphy,
      D(0), D(1), D(2), D(3), D(4), D(5), D(6), D(7), D(8), D(9), D(10), D(11), D(12), D(13), D(14), D(15),
      D(16), D(17), D(18), D(19), D(20), D(21), D(22), D(23), D(24), D(25), (int *)p[26], D(27), D(28), D(29),
      D(30), D(31), D(32), D(33), D(34), D(35), D(36), D(37), D(38), D(39), D(40), D(41), D(42), D(43), D(44),
      D(45), D(46), D(47), D(48), D(49), D(50), D(51), D(52), D(53), D(54), D(55));
#undef D
  }
  return omp_get_wtime() - t0;
}
