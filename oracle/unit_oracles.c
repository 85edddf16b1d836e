/*
 * unit_oracles.c -- per-process CPU restatements of three pieces of the reference kernel.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * The whole-column oracles (oracle/_ref, cloudsc_oracle.c) pin the scheme end to end; these isolate the three processes
 * whose device implementation is organised differently from the reference, so that a unit test can feed them adversarial
 * inputs the 100-column data set never produces:
 *
 *   unit_sink_limiter()    src/cloudsc_c/cloudsc/cloudsc_c.c:2233-2286  on the DENSE 5x5 ZSOLQA (the device code keeps the
 *                          antisymmetric upper triangle + diagonal and tests all species up front)
 *   unit_implicit_solve()  :2294-2412  dense LHS, generic non-pivoting LU, back-substitution, clipping (the device code
 *                          uses the closed form for the four non-zero ZSOLQB entries)
 *   unit_satadj_step()     :1162-1168 (= :1169-1175)  one Newton step of the saturation adjustment with the statement
 *                          functions FOEEWM / FOEDEM / FOELDCPM inlined as the reference has them (the device code
 *                          shares the exponentials and reciprocals and, in its fast mode, folds two divisions into one)
 *
 * Statement order and association follow the reference lines cited; compile with FP contraction off.
 * -> oracle/libcloudsc_unit.so (make -C oracle unit)
 */
#include <math.h>

#define N 5

/* :2233-2286.  zsolqa[n][m]: source of species m from species n, as in the reference (first index = "from"). */
void unit_sink_limiter(double zsolqa[N][N], const double zqx[N], double zepsec) {
  for (int jm = 0; jm < N; jm++) {
    double psum = 0.0;
    for (int jn = 0; jn < N; jn++) psum = psum + zsolqa[jn][jm];     /* :2252-2257 */
    const double zsinksum = 0.0 - psum;                             /* :2259-2262 */
    const double zmm = fmax(zqx[jm], zepsec);                       /* :2267-2271 */
    const double zrr = fmax(zsinksum, zmm);
    const double zzratio = zmm / zrr;
    for (int jn = 0; jn < N; jn++)                                   /* :2276-2286 */
      if (zsolqa[jn][jm] < 0.0) {
        zsolqa[jn][jm] = zsolqa[jn][jm] * zzratio;
        zsolqa[jm][jn] = zsolqa[jm][jn] * zzratio;
      }
  }
}

/* :2294-2412.  zsolqb[m][n] as in the reference; zfallsink[m]; returns zqxn[5]. */
void unit_implicit_solve(const double zsolqa[N][N], const double zsolqb[N][N], const double zqx[N], const double zfallsink[N],
                         double zepsec, double zqxn[N]) {
  double zqlhs[N][N];
  for (int jm = 0; jm < N; jm++)
    for (int jn = 0; jn < N; jn++) {
      if (jn == jm) {
        zqlhs[jm][jn] = 1.0 + zfallsink[jm];                        /* :2302-2306 */
        for (int jo = 0; jo < N; jo++) zqlhs[jm][jn] = zqlhs[jm][jn] + zsolqb[jn][jo];
      } else {
        zqlhs[jm][jn] = -zsolqb[jm][jn];                            /* :2313-2315 */
      }
    }
  for (int jm = 0; jm < N; jm++) {                                   /* :2326-2337 */
    double zexplicit = 0.0;
    for (int jn = 0; jn < N; jn++) zexplicit = zexplicit + zsolqa[jn][jm];
    zqxn[jm] = zqx[jm] + zexplicit;
  }
  for (int jn = 0; jn < N - 1; jn++)                                 /* :2351-2364 non-pivoting factorisation */
    for (int jm = jn + 1; jm < N; jm++) {
      zqlhs[jn][jm] = zqlhs[jn][jm] / zqlhs[jn][jn];
      for (int ik = jn + 1; ik < N; ik++) zqlhs[ik][jm] = zqlhs[ik][jm] - zqlhs[jn][jm] * zqlhs[ik][jn];
    }
  for (int jn = 1; jn < N; jn++)                                     /* :2369-2377 */
    for (int jm = 0; jm < jn; jm++) zqxn[jn] = zqxn[jn] - zqlhs[jm][jn] * zqxn[jm];
  zqxn[N - 1] = zqxn[N - 1] / zqlhs[N - 1][N - 1];                   /* :2380-2382 */
  for (int jn = N - 2; jn >= 0; jn--) {                              /* :2384-2397 */
    for (int jm = jn + 1; jm < N; jm++) zqxn[jn] = zqxn[jn] - zqlhs[jm][jn] * zqxn[jm];
    zqxn[jn] = zqxn[jn] / zqlhs[jn][jn];
  }
  for (int jn = 0; jn < N - 1; jn++)                                 /* :2402-2412 */
    if (zqxn[jn] < zepsec) {
      zqxn[N - 1] = zqxn[N - 1] + zqxn[jn];
      zqxn[jn] = 0.0;
    }
}

/* :1162-1168: ztp1 and zqsmix after one Newton step; k[] = rtt r2es r3les r3ies r4les r4ies r5alvcp r5alscp ralvdcp ralsdcp
 * rtwat rtice rtwat_rtice_r retv; zqp = 1/pap */
void unit_satadj_step(const double k[14], double zqp, double *ztp1, double *zqsmix) {
  const double rtt = k[0], r2es = k[1], r3les = k[2], r3ies = k[3], r4les = k[4], r4ies = k[5], r5alvcp = k[6], r5alscp = k[7],
               ralvdcp = k[8], ralsdcp = k[9], rtwat = k[10], rtice = k[11], rtwat_rtice_r = k[12], retv = k[13];
  const double t = *ztp1;
#define FOEALFA fmin(1.0, pow((fmax(rtice, fmin(rtwat, t)) - rtice) * rtwat_rtice_r, 2))
  const double foeewm = r2es * (FOEALFA * exp((r3les * (t - rtt)) / (t - r4les)) + (1.0 - FOEALFA) * exp((r3ies * (t - rtt)) / (t - r4ies)));
  double zqsat = foeewm * zqp;                                      /* :1162 */
  zqsat = fmin(0.5, zqsat);                                         /* :1163 */
  const double zcor = 1.0 / (1.0 - retv * zqsat);                   /* :1164 */
  zqsat = zqsat * zcor;                                             /* :1165 */
  const double foedem = (FOEALFA * r5alvcp) * (1.0 / pow(t - r4les, 2)) + ((1.0 - FOEALFA) * r5alscp) * (1.0 / pow(t - r4ies, 2));
  const double zcond = (*zqsmix - zqsat) / (1.0 + (zqsat * zcor) * foedem);   /* :1166 */
  const double foeldcpm = FOEALFA * ralvdcp + (1.0 - FOEALFA) * ralsdcp;
  *ztp1 = t + foeldcpm * zcond;                                     /* :1167 */
  *zqsmix = *zqsmix - zcond;                                        /* :1168 */
#undef FOEALFA
}
