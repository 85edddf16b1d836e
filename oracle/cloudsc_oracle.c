/*
 * cloudsc_oracle.c -- CPU restatement of the CLOUDSC kernel.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may
 * build, load or call this file; the product library has no CPU path.
 *
 * What it restates: the algorithm of src/cloudsc_c/cloudsc/cloudsc_c.c:360-2582 of the reference
 * (dwarf-cloudsc-c), one column at a time in plain scalar C: per-column arrays over the klev levels,
 * the same multi-pass structure as the reference (initial state -> tidy-up -> saturation -> sweep
 * from ncldtop to klev -> flux diagnostics) and the same GENERIC 5x5 non-pivoting LU solve -- i.e.
 * deliberately NOT the k-cached, sparse-solve organisation of the CUDA kernel, so that the two are
 * independent implementations of the same arithmetic.  Each section cites the reference lines it
 * follows.  Dead code of the reference (ztrpaus :697-712, the first sink-sum pass :2207-2227, the
 * switched-off process variants iwarmrain=1 / ievaprain=1 / ievapsnow=2 / idepice=2 and, with them, the
 * laerliqautolsp / laerliqcoll branches that only exist inside iwarmrain=1) is not restated; the two live aerosol
 * switches laericesed (:1561-1566) and laericeauto (:1625-1632) are; oracle_check_params() below states the parameter domain.
 *
 * PARITY PINNING: checked by tests/test_cpu_oracle.py against (a) the golden vectors of the
 * reference's own data set, tests/golden/reference.h5 (21 fields x 100 columns, produced by the
 * Fortran kernel; bar L1-rel <= 1e-15, point-wise <= 1e-12) and (b) BIT-FOR-BIT against the
 * unmodified reference kernel compiled from /root/reference by oracle/Makefile (oracle/_ref), on the
 * shipped columns and on eight perturbed input sets that reach the warm-phase / land / detrainment /
 * rain-freezing branches the shipped data never takes.  Compile with FP contraction off.
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include "cloudsc_params.h"

#define NCLV 5
enum { QL = 0, QI = 1, QR = 2, QS = 3, QV = 4 };

#define MAXLEV 400

static double sq2(double x) { return x * x; }
/* statement functions, cloudsc_c.c:332-351 */
static double f_foealfa(double t) { return fmin(1.0, sq2((fmax(rtice, fmin(rtwat, t)) - rtice) * rtwat_rtice_r)); }
static double f_eliq(double t) { return exp((r3les * (t - rtt)) / (t - r4les)); }
static double f_eice(double t) { return exp((r3ies * (t - rtt)) / (t - r4ies)); }
static double f_foeewm(double t) { return r2es * (f_foealfa(t) * f_eliq(t) + (1.0 - f_foealfa(t)) * f_eice(t)); }
static double f_foedem(double t) {
  return (f_foealfa(t) * r5alvcp) * (1.0 / sq2(t - r4les)) + ((1.0 - f_foealfa(t)) * r5alscp) * (1.0 / sq2(t - r4ies));
}
static double f_foeldcpm(double t) { return f_foealfa(t) * ralvdcp + (1.0 - f_foealfa(t)) * ralsdcp; }

/* 1 if the parameter set is inside what this restatement (and the CUDA kernel) covers */
int oracle_check_params(void) {
  const struct TECLDP *y = yrecldp;
  if (!y) return 0;
  if (y->laerliqautolsp || y->laerliqcoll) return 0;
  if (y->nssopt < 0 || y->nssopt > 3 || y->ncldtop < 2) return 0;
  return 1;
}

struct colptr {            /* one column of every field: base pointer + level stride (= klon) */
  const double *pt, *pq, *tt, *tq, *ta, *tcld, *pvfl, *pvfi, *phrsw, *phrlw, *pvervel, *pap, *paph,
      *plu, *psnde, *pmfu, *pmfd, *pa, *pclv, *psupsat, *pre_ice, *picrit_aer, *pnice;
  double plsm;
  int ktype;
  double *plude, *lt, *lq, *la, *lcld, *pcovptot, *prainfrac;
  double *fsqlf, *fsqif, *fcqnng, *fcqlng, *fsqrf, *fsqsf, *fcqrng, *fcqsng, *fsqltur, *fsqitur, *fplsl,
      *fplsn, *fhpsl, *fhpsn;
  size_t ls;               /* level stride */
  size_t ss;               /* species stride = klev*klon */
};

static void column(const struct colptr *c, const int klev, const double ptsphy) {
  const struct TECLDP *y = yrecldp;
  const size_t ls = c->ls, ss = c->ss;
#define AT(p, k) ((p)[(size_t)(k) * ls])
#define AT3(p, m, k) ((p)[(size_t)(m) * ss + (size_t)(k) * ls])

  static const int iphase[NCLV] = {1, 2, 1, 2, 0};   /* :400-404 */
  static const int imelt[NCLV] = {1, 2, 3, 2, -1};   /* :408-412, 0-based target species */
  const double zepsilon = 100.0 * DBL_EPSILON;       /* :360 */
  const double zqtmst = 1.0 / ptsphy;                /* :385-391 */
  const double zrdcp = rd / rcpd;
  const double zepsec = 1.0e-14;
  const double zrg_r = 1.0 / rg;
  const double zrldcp = 1.0 / (ralsdcp - ralvdcp);
  const double ztw1 = 1329.31, ztw2 = 0.0074615, ztw3 = 0.85e5, ztw4 = 40.637, ztw5 = 275.0;

  double zvqx[NCLV];                                 /* fall speeds :437-456 */
  int llfall[NCLV];
  zvqx[QV] = 0.0; zvqx[QL] = 0.0; zvqx[QI] = y->rvice; zvqx[QR] = y->rvrain; zvqx[QS] = y->rvsnow;
  for (int m = 0; m < NCLV; m++) llfall[m] = zvqx[m] > 0.0;
  llfall[QI] = 0;

  /* ---- per-level arrays of this column (the reference's [klev][klon] temporaries) ---------- */
  double ztp1[MAXLEV], za[MAXLEV], zaorig[MAXLEV], zli[MAXLEV], zliqfrac[MAXLEV], zicefrac[MAXLEV];
  double zqx[NCLV][MAXLEV], zqx0[NCLV][MAXLEV], zlneg[NCLV][MAXLEV], zqxn2d[NCLV][MAXLEV];
  double zpfplsx[NCLV][MAXLEV + 1];
  double zfoealfa[MAXLEV], zfoeewmt[MAXLEV], zqsmix[MAXLEV], zfoeew[MAXLEV], zqsice[MAXLEV],
      zfoeeliqt[MAXLEV], zqsliq[MAXLEV];
  double tlt[MAXLEV], tlq[MAXLEV], tla[MAXLEV], tlcld[4][MAXLEV];

  /* ---- 1. initial values, :415-513 ---------------------------------------------------------- */
  for (int k = 0; k < klev; k++) {
    tlt[k] = tlq[k] = tla[k] = 0.0;
    for (int m = 0; m < 4; m++) tlcld[m][k] = 0.0;
    ztp1[k] = AT(c->pt, k) + ptsphy * AT(c->tt, k);
    zqx[QV][k] = zqx0[QV][k] = AT(c->pq, k) + ptsphy * AT(c->tq, k);
    za[k] = zaorig[k] = AT(c->pa, k) + ptsphy * AT(c->ta, k);
    for (int m = 0; m < 4; m++) zqx[m][k] = zqx0[m][k] = AT3(c->pclv, m, k) + ptsphy * AT3(c->tcld, m, k);
    for (int m = 0; m < NCLV; m++) { zqxn2d[m][k] = 0.0; zlneg[m][k] = 0.0; }
  }
  for (int m = 0; m < NCLV; m++)
    for (int k = 0; k <= klev; k++) zpfplsx[m][k] = 0.0;
  double rainfrac_toprfz = 0.0;

  /* ---- tidy up, :519-575 --------------------------------------------------------------------- */
  for (int k = 0; k < klev; k++) {
    if (zqx[QL][k] + zqx[QI][k] < y->rlmin || za[k] < y->ramin) {
      zlneg[QL][k] += zqx[QL][k];
      double zqadj = zqx[QL][k] * zqtmst;
      tlq[k] += zqadj;
      tlt[k] -= ralvdcp * zqadj;
      zqx[QV][k] += zqx[QL][k];
      zqx[QL][k] = 0.0;
      zlneg[QI][k] += zqx[QI][k];
      zqadj = zqx[QI][k] * zqtmst;
      tlq[k] += zqadj;
      tlt[k] -= ralsdcp * zqadj;
      zqx[QV][k] += zqx[QI][k];
      zqx[QI][k] = 0.0;
      za[k] = 0.0;
    }
  }
  for (int m = 0; m < 4; m++)
    for (int k = 0; k < klev; k++)
      if (zqx[m][k] < y->rlmin) {
        zlneg[m][k] += zqx[m][k];
        const double zqadj = zqx[m][k] * zqtmst;
        tlq[k] += zqadj;
        if (iphase[m] == 1) tlt[k] -= ralvdcp * zqadj;
        if (iphase[m] == 2) tlt[k] -= ralsdcp * zqadj;
        zqx[QV][k] += zqx[m][k];
        zqx[m][k] = 0.0;
      }

  /* ---- saturation values, :583-640 ---------------------------------------------------------- */
  for (int k = 0; k < klev; k++) {
    const double t = ztp1[k], p = AT(c->pap, k);
    zfoealfa[k] = f_foealfa(t);
    zfoeewmt[k] = fmin(f_foeewm(t) / p, 0.5);
    zqsmix[k] = zfoeewmt[k] / (1.0 - retv * zfoeewmt[k]);
    const double zalfa = fmax(0.0, copysign(1.0, t - rtt));
    zfoeew[k] = fmin((zalfa * (r2es * f_eliq(t)) + (1.0 - zalfa) * (r2es * f_eice(t))) / p, 0.5);
    zfoeew[k] = fmin(0.5, zfoeew[k]);
    zqsice[k] = zfoeew[k] / (1.0 - retv * zfoeew[k]);
    zfoeeliqt[k] = fmin((r2es * f_eliq(t)) / p, 0.5);
    zqsliq[k] = zfoeeliqt[k] / (1.0 - retv * zfoeeliqt[k]);
  }
  for (int k = 0; k < klev; k++) {
    za[k] = fmax(0.0, fmin(1.0, za[k]));
    zli[k] = zqx[QL][k] + zqx[QI][k];
    if (zli[k] > y->rlmin) { zliqfrac[k] = zqx[QL][k] / zli[k]; zicefrac[k] = 1.0 - zliqfrac[k]; }
    else { zliqfrac[k] = 0.0; zicefrac[k] = 0.0; }
  }

  /* ---- single-level carries, :717-724 ------------------------------------------------------- */
  double zanewm1 = 0.0, zcovpclr = 0.0, zcovpmax = 0.0, zcovptot = 0.0, zcldtopdist = 0.0;
  double zqxnm1[NCLV] = {0, 0, 0, 0, 0};
  const double paph_surf = AT(c->paph, klev);

  /* ============================ 3. vertical sweep, :732-2508 ================================= */
  for (int k = y->ncldtop - 1; k < klev; k++) {
    const int jk = k + 1;                       /* the reference's 1-based level */
    const double t = ztp1[k], pap = AT(c->pap, k);
    double zqxfg[NCLV], zsolqa[NCLV][NCLV], zsolqb[NCLV][NCLV], zqlhs[NCLV][NCLV], zqxn[NCLV];
    double zfallsrce[NCLV], zfallsink[NCLV], zconvsrce[NCLV], zconvsink[NCLV], zpsupsatsrce[NCLV];
    for (int m = 0; m < NCLV; m++) {            /* :739-797 */
      zqxfg[m] = zqx[m][k];
      zfallsrce[m] = zfallsink[m] = zconvsrce[m] = zconvsink[m] = zpsupsatsrce[m] = 0.0;
      for (int n = 0; n < NCLV; n++) { zsolqa[m][n] = 0.0; zsolqb[m][n] = 0.0; }
    }
    double zrainaut = 0.0, zrainacc = 0.0, zsnowaut, zacust = 0.0, zqpretot = 0.0, zlfinalsum = 0.0;
    double zlcond1 = 0.0, zlcond2 = 0.0, zsupsat = 0.0, zsolab = 0.0, zsolac = 0.0, zldefr = 0.0;
    (void)zrainaut; (void)zrainacc;

    /* derived variables, :799-841 */
    const double zdp = AT(c->paph, k + 1) - AT(c->paph, k);
    const double zgdp = rg / zdp;
    const double zrho = pap / (rd * t);
    const double zdtgdp = ptsphy * zgdp;
    const double zrdtgdp = zdp * (1.0 / (ptsphy * rg));
    const double zfacw = r5les / sq2(t - r4les);
    double zcor = 1.0 / (1.0 - retv * zfoeeliqt[k]);
    const double zdqsliqdt = (zfacw * zcor) * zqsliq[k];
    const double zcorqsliq = 1.0 + ralvdcp * zdqsliqdt;
    const double zfaci0 = r5ies / sq2(t - r4ies);
    zcor = 1.0 / (1.0 - retv * zfoeew[k]);
    const double zdqsicedt = (zfaci0 * zcor) * zqsice[k];
    const double zcorqsice = 1.0 + ralsdcp * zdqsicedt;
    double zalfaw = zfoealfa[k];
    double zfac = zalfaw * zfacw + (1.0 - zalfaw) * zfaci0;
    zcor = 1.0 / (1.0 - retv * zfoeewmt[k]);
    const double zdqsmixdt = (zfac * zcor) * zqsmix[k];
    const double zcorqsmix = 1.0 + f_foeldcpm(t) * zdqsmixdt;
    const double zevaplimmix = fmax((zqsmix[k] - zqx[QV][k]) / zcorqsmix, 0.0);
    (void)zcorqsliq;
    double ztmpa = 1.0 / fmax(za[k], zepsec);
    double zliqcld = zqx[QL][k] * ztmpa, zicecld = zqx[QI][k] * ztmpa;
    double zlicld = zliqcld + zicecld;

    /* evaporate very small amounts of liquid and ice, :846-859 */
    if (zqx[QL][k] < y->rlmin) { zsolqa[QL][QV] = zqx[QL][k]; zsolqa[QV][QL] = -zqx[QL][k]; }
    if (zqx[QI][k] < y->rlmin) { zsolqa[QI][QV] = zqx[QI][k]; zsolqa[QV][QI] = -zqx[QI][k]; }

    /* 3.1 ice supersaturation adjustment, :874-954 */
    const double zfokoop = fmin(rkoop1 - rkoop2 * t, (r2es * f_eliq(t)) / (r2es * f_eice(t)));
    double zfaci;
    if (y->nssopt == 0 || t >= rtt) { zfac = 1.0; zfaci = 1.0; }
    else { zfac = za[k] + zfokoop * (1.0 - za[k]); zfaci = ptsphy / y->rkooptau; }
    if (za[k] > 1.0 - y->ramin) {
      zsupsat = fmax((zqx[QV][k] - zfac * zqsice[k]) / zcorqsice, 0.0);
    } else {
      const double zqp1env = (zqx[QV][k] - za[k] * zqsice[k]) / fmax(1.0 - za[k], zepsilon);
      zsupsat = fmax(((1.0 - za[k]) * (zqp1env - zfac * zqsice[k])) / zcorqsice, 0.0);
    }
    if (zsupsat > zepsec) {
      const int to = t > y->rthomo ? QL : QI;
      zsolqa[QV][to] += zsupsat;
      zsolqa[to][QV] -= zsupsat;
      zqxfg[to] += zsupsat;
      zsolac = (1.0 - za[k]) * zfaci;
    }
    if (AT(c->psupsat, k) > zepsec) {
      const int to = t > y->rthomo ? QL : QI;
      zsolqa[to][to] += AT(c->psupsat, k);
      zpsupsatsrce[to] = AT(c->psupsat, k);
      zqxfg[to] += AT(c->psupsat, k);
      zsolac = (1.0 - za[k]) * zfaci;
    }

    /* 3.2 detrainment from convection, :967-987 (plude is scaled IN PLACE) */
    if (jk < klev) {
      AT(c->plude, k) = AT(c->plude, k) * zdtgdp;
      if (AT(c->plu, k + 1) > zepsec && AT(c->plude, k) > y->rlmin) {
        zsolac += AT(c->plude, k) / AT(c->plu, k + 1);
        zalfaw = zfoealfa[k];
        zconvsrce[QL] = zalfaw * AT(c->plude, k);
        zconvsrce[QI] = (1.0 - zalfaw) * AT(c->plude, k);
        zsolqa[QL][QL] += zconvsrce[QL];
        zsolqa[QI][QI] += zconvsrce[QI];
      } else {
        AT(c->plude, k) = 0.0;
      }
      zsolqa[QS][QS] += AT(c->psnde, k) * zdtgdp;
    }

    /* 3.3 subsidence compensating convective updraughts, :1002-1075 */
    if (jk > y->ncldtop) {
      const double zmf = fmax(0.0, (AT(c->pmfu, k) + AT(c->pmfd, k)) * zdtgdp);
      zacust = zmf * zanewm1;
      double zlcust[NCLV];
      for (int m = 0; m < NCLV; m++)
        if (!llfall[m] && iphase[m] > 0) { zlcust[m] = zmf * zqxnm1[m]; zconvsrce[m] += zlcust[m]; }
      const double zdtdp = ((zrdcp * 0.5) * (ztp1[k - 1] + t)) / AT(c->paph, k);
      const double zdtforc = zdtdp * (pap - AT(c->pap, k - 1));
      const double zdqs = (zanewm1 * zdtforc) * zdqsmixdt;
      for (int m = 0; m < NCLV; m++)
        if (!llfall[m] && iphase[m] > 0) {
          double zlfinal = fmax(0.0, zlcust[m] - zdqs);
          const double zevap = fmin(zlcust[m] - zlfinal, zevaplimmix);
          zlfinal = zlcust[m] - zevap;
          zlfinalsum += zlfinal;
          zsolqa[m][m] += zlcust[m];
          zsolqa[m][QV] += zevap;
          zsolqa[QV][m] -= zevap;
        }
      if (zlfinalsum < zepsec) zacust = 0.0;
      zsolac += zacust;
    }
    if (jk < klev) {
      const double zmfdn = fmax(0.0, (AT(c->pmfu, k + 1) + AT(c->pmfd, k + 1)) * zdtgdp);
      zsolab += zmfdn;
      zsolqb[QL][QL] += zmfdn;
      zsolqb[QI][QI] += zmfdn;
      zconvsink[QL] = zmfdn;
      zconvsink[QI] = zmfdn;
    }

    /* 3.4 erosion of clouds by turbulent mixing, :1087-1118 */
    double zldifdt = y->rcldiff * ptsphy;
    if (c->ktype > 0 && AT(c->plude, k) > zepsec) zldifdt = y->rcldiff_convi * zldifdt;
    if (zli[k] > zepsec) {
      const double ze = zldifdt * fmax(zqsmix[k] - zqx[QV][k], 0.0);
      double zleros = za[k] * ze;
      zleros = fmin(zleros, zevaplimmix);
      zleros = fmin(zleros, zli[k]);
      const double zaeros = zleros / zlicld;
      zsolac -= zaeros;
      zsolqa[QL][QV] += zliqfrac[k] * zleros;
      zsolqa[QV][QL] -= zliqfrac[k] * zleros;
      zsolqa[QI][QV] += zicefrac[k] * zleros;
      zsolqa[QV][QI] -= zicefrac[k] * zleros;
    }

    /* 3.4 condensation/evaporation due to dqsat/dT, :1137-1182 */
    double zdqs;
    {
      const double zdtdp = (zrdcp * t) / pap;
      const double zdpmxdt = zdp * zqtmst;
      double zmfdn = 0.0;
      if (jk < klev) zmfdn = AT(c->pmfu, k + 1) + AT(c->pmfd, k + 1);
      double zwtot = AT(c->pvervel, k) + (0.5 * rg) * (AT(c->pmfu, k) + AT(c->pmfd, k) + zmfdn);
      zwtot = fmin(zdpmxdt, fmax(-zdpmxdt, zwtot));
      const double zzzdt = AT(c->phrsw, k) + AT(c->phrlw, k);
      const double zdtdiab = fmin(zdpmxdt * zdtdp, fmax(-zdpmxdt * zdtdp, zzzdt)) * ptsphy + ralfdcp * zldefr;
      const double zdtforc = (zdtdp * zwtot) * ptsphy + zdtdiab;
      const double zqold = zqsmix[k];
      double tn = fmax(t + zdtforc, 160.0), qsm = zqsmix[k];
      const double zqp = 1.0 / pap;
      for (int it = 0; it < 2; it++) {          /* two Newton steps, :1162-1175 */
        double zqsat = fmin(0.5, f_foeewm(tn) * zqp);
        zcor = 1.0 / (1.0 - retv * zqsat);
        zqsat = zqsat * zcor;
        const double zcond = (qsm - zqsat) / (1.0 + (zqsat * zcor) * f_foedem(tn));
        tn = tn + f_foeldcpm(tn) * zcond;
        qsm = qsm - zcond;
      }
      zdqs = qsm - zqold;
    }
    /* 3.4a evaporation of clouds, :1189-1207 */
    if (zdqs > 0.0) {
      double zlevap = za[k] * fmin(zdqs, zlicld);
      zlevap = fmin(zlevap, zevaplimmix);
      zlevap = fmin(zlevap, fmax(zqsmix[k] - zqx[QV][k], 0.0));
      zsolqa[QL][QV] += zliqfrac[k] * zlevap;
      zsolqa[QV][QL] -= zliqfrac[k] * zlevap;
      zsolqa[QI][QV] += zicefrac[k] * zlevap;
      zsolqa[QV][QI] -= zicefrac[k] * zlevap;
    }
    /* 3.4b (1) more cloud water in existing clouds, :1213-1250 */
    if (zdqs <= -y->rlmin && za[k] > zepsec) {
      zlcond1 = fmax(-zdqs, 0.0);
      double zcdmax;
      if (za[k] > 0.99) {
        zcor = 1.0 / (1.0 - retv * zqsmix[k]);
        zcdmax = (zqx[QV][k] - zqsmix[k]) / (1.0 + (zcor * zqsmix[k]) * f_foedem(t));
      } else {
        zcdmax = (zqx[QV][k] - za[k] * zqsmix[k]) / za[k];
      }
      zlcond1 = fmax(fmin(zlcond1, zcdmax), 0.0);
      zlcond1 = za[k] * zlcond1;
      if (zlcond1 < y->rlmin) zlcond1 = 0.0;
      const int to = t > y->rthomo ? QL : QI;
      zsolqa[QV][to] += zlcond1;
      zsolqa[to][QV] -= zlcond1;
      zqxfg[to] += zlcond1;
    }
    /* 3.4b (2) generation of new clouds, :1253-1363 */
    if (zdqs <= -y->rlmin && za[k] < 1.0 - zepsec) {
      double zrhc = y->ramid;
      const double zsigk = pap / paph_surf;
      if (zsigk > 0.8) zrhc = y->ramid + (1.0 - y->ramid) * sq2((zsigk - 0.8) / 0.2);
      double zqe = 0.0;
      if (y->nssopt == 0 || y->nssopt == 1) {
        zqe = (zqx[QV][k] - za[k] * zqsice[k]) / fmax(zepsec, 1.0 - za[k]);
        zqe = fmax(0.0, zqe);
      } else if (y->nssopt == 2) {
        zqe = zqx[QV][k];
      } else if (y->nssopt == 3) {
        zqe = zqx[QV][k] + zli[k];
      }
      zfac = (y->nssopt == 0 || t >= rtt) ? 1.0 : zfokoop;
      if (zqe >= zqsice[k] * zfac * zrhc && zqe < zqsice[k] * zfac) {
        double zacond = (-((1.0 - za[k]) * zfac) * zdqs) / fmax(2.0 * (zfac * zqsice[k] - zqe), zepsec);
        zacond = fmin(zacond, 1.0 - za[k]);
        zlcond2 = (-(zfac * zdqs) * 0.5) * zacond;
        const double zzdl = (2.0 * (zfac * zqsice[k] - zqe)) / fmax(zepsec, 1.0 - za[k]);
        if (zdqs * zfac < -zzdl) {
          const double zlcondlim = ((za[k] - 1.0) * zfac) * zdqs - zfac * zqsice[k] + zqx[QV][k];
          zlcond2 = fmin(zlcond2, zlcondlim);
        }
        zlcond2 = fmax(zlcond2, 0.0);
        if (1.0 - za[k] < zepsec || zlcond2 < y->rlmin) { zlcond2 = 0.0; zacond = 0.0; }
        if (zlcond2 == 0.0) zacond = 0.0;
        zsolac += zacond;
        const int to = t > y->rthomo ? QL : QI;
        zsolqa[QV][to] += zlcond2;
        zsolqa[to][QV] -= zlcond2;
        zqxfg[to] += zlcond2;
      }
    }

    /* 3.7 growth of ice by vapour deposition (Rotstayn 2001), :1382-1447 */
    if (za[k] >= y->rcldtopcf && za[k - 1] < y->rcldtopcf) zcldtopdist = 0.0;
    else zcldtopdist = zcldtopdist + zdp / (zrho * rg);
    if (zqxfg[QL] > y->rlmin && t < rtt) {
      const double zvpice = ((r2es * f_eice(t)) * rv) / rd;
      const double zvpliq = zvpice * zfokoop;
      const double zicenuclei = 1000.0 * exp((12.96 * (zvpliq - zvpice)) / zvpliq - 0.639);
      const double zadd = (rlstt * (rlstt / (rv * t) - 1.0)) / (0.024 * t);
      const double zbdd = ((rv * t) * pap) / (2.21 * zvpice);
      const double zcvds = ((7.8 * pow(zicenuclei / zrho, 0.666)) * (zvpliq - zvpice)) / ((8.87 * (zadd + zbdd)) * zvpice);
      const double zice0 = fmax(zicecld, (zicenuclei * y->riceinit) / zrho);
      const double zinew = pow((0.666 * zcvds) * ptsphy + pow(zice0, 0.666), 1.5);
      double zdepos = fmax(za[k] * (zinew - zice0), 0.0);
      zdepos = fmin(zdepos, zqxfg[QL]);
      const double zinfactor = fmin(zicenuclei / 15000.0, 1.0);
      zdepos = zdepos * fmin(zinfactor + (1.0 - zinfactor) * (y->rdepliqrefrate + zcldtopdist / y->rdepliqrefdepth), 1.0);
      zsolqa[QL][QI] += zdepos;
      zsolqa[QI][QL] -= zdepos;
      zqxfg[QI] += zdepos;
      zqxfg[QL] -= zdepos;
    }

    /* 4. precipitation processes: revised in-cloud condensate, :1528-1533 */
    ztmpa = 1.0 / fmax(za[k], zepsec);
    zliqcld = zqxfg[QL] * ztmpa;
    zicecld = zqxfg[QI] * ztmpa;
    zlicld = zliqcld + zicecld;

    /* 4.2 sedimentation, :1541-1576 */
    for (int m = 0; m < NCLV; m++)
      if (llfall[m] || m == QI) {
        if (jk > y->ncldtop) {
          zfallsrce[m] = zpfplsx[m][k] * zdtgdp;
          zsolqa[m][m] += zfallsrce[m];
          zqxfg[m] += zfallsrce[m];
          zqpretot += zqxfg[m];
        }
        /* aerosol override of the ice fall speed, :1561-1566 (the reference stores it in the shared zvqx[1] and reads
           it back on the next line for the same column; nothing else reads zvqx during the sweep) */
        double zv = zvqx[m];
        if (y->laericesed && m == QI) zv = 0.002 * pow(AT(c->pre_ice, k), 1.0);
        const double zfall = zv * zrho;
        zfallsink[m] = zdtgdp * zfall;
      }
    /* precipitation cover, :1594-1611 */
    double zraincld, zsnowcld;
    if (zqpretot > zepsec) {
      zcovptot = 1.0 - ((1.0 - zcovptot) * (1.0 - fmax(za[k], za[k - 1]))) / (1.0 - fmin(za[k - 1], 1.0 - 1.0e-6));
      zcovptot = fmax(zcovptot, y->rcovpmin);
      zcovpclr = fmax(0.0, zcovptot - za[k]);
      zraincld = zqxfg[QR] / zcovptot;
      zsnowcld = zqxfg[QS] / zcovptot;
      zcovpmax = fmax(zcovptot, zcovpmax);
    } else {
      zraincld = 0.0; zsnowcld = 0.0; zcovptot = 0.0; zcovpclr = 0.0; zcovpmax = 0.0;
    }

    /* 4.3a autoconversion to snow, :1616-1637 */
    if (t <= rtt && zicecld > zepsec) {
      double zzco = (ptsphy * y->rsnowlin1) * exp(y->rsnowlin2 * (t - rtt));
      double zlcrit = y->rlcritsnow;
      if (y->laericeauto) {                     /* :1625-1632 */
        zlcrit = AT(c->picrit_aer, k);
        zzco = zzco * pow(y->rnice / AT(c->pnice, k), 0.333);
      }
      zsnowaut = zzco * (1.0 - exp(-sq2(zicecld / zlcrit)));
      zsolqb[QI][QS] += zsnowaut;
    }
    /* 4.3b warm rain, Khairoutdinov & Kogan (iwarmrain = 2), :1644-1759 */
    if (zliqcld > zepsec) {
      double zconst, zlcrit;
      if (c->plsm > 0.5) { zconst = y->rcl_kk_cloud_num_land; zlcrit = y->rclcrit_land; }
      else { zconst = y->rcl_kk_cloud_num_sea; zlcrit = y->rclcrit_sea; }
      if (zliqcld > zlcrit) {
        zrainaut = ((((1.5 * za[k]) * ptsphy) * y->rcl_kkaau) * pow(zliqcld, y->rcl_kkbauq)) * pow(zconst, y->rcl_kkbaun);
        zrainaut = fmin(zrainaut, zqxfg[QL]);
        if (zrainaut < zepsec) zrainaut = 0.0;
        zrainacc = (((2.0 * za[k]) * ptsphy) * y->rcl_kkaac) * pow(zliqcld * zraincld, y->rcl_kkbac);
        zrainacc = fmin(zrainacc, zqxfg[QL]);
        if (zrainacc < zepsec) zrainacc = 0.0;
      } else {
        zrainaut = 0.0; zrainacc = 0.0;
      }
      const int to = t <= rtt ? QS : QR;
      zsolqa[QL][to] += zrainaut;
      zsolqa[QL][to] += zrainacc;
      zsolqa[to][QL] -= zrainaut;
      zsolqa[to][QL] -= zrainacc;
    }
    /* riming of snow by cloud water, :1768-1808 */
    if (t <= rtt && zliqcld > zepsec) {
      const double zfallcorr = pow(y->rdensref / zrho, 0.4);
      if (zcovptot > 0.01 && zsnowcld > zepsec) {
        double zsnowrime = ((((0.3 * zcovptot) * ptsphy) * y->rcl_const7s) * zfallcorr) *
                           pow((zrho * zsnowcld) * y->rcl_const1s, y->rcl_const8s);
        zsnowrime = fmin(zsnowrime, 1.0);
        zsolqb[QL][QS] += zsnowrime;
      }
    }

    /* 4.4a melting of snow and ice, :1817-1859 */
    const double zicetot = zqxfg[QI] + zqxfg[QS];
    double zmeltmax = 0.0;
    if (zicetot > zepsec && t > rtt) {
      const double zsubsat = fmax(zqsice[k] - zqx[QV][k], 0.0);
      const double ztdmtw0 = t - rtt - zsubsat * (ztw1 + ztw2 * (pap - ztw3) - ztw4 * (t - ztw5));
      const double zcons1 = fabs((ptsphy * (1.0 + 0.5 * ztdmtw0)) / y->rtaumel);
      zmeltmax = fmax((ztdmtw0 * zcons1) * zrldcp, 0.0);
    }
    for (int m = 0; m < NCLV; m++)
      if (iphase[m] == 2) {
        const int n = imelt[m];
        if (zicetot > zepsec && zmeltmax > zepsec) {
          const double zalfa = zqxfg[m] / zicetot;
          const double zmelt = fmin(zqxfg[m], zalfa * zmeltmax);
          zqxfg[m] -= zmelt;
          zqxfg[n] += zmelt;
          zsolqa[m][n] += zmelt;
          zsolqa[n][m] -= zmelt;
        }
      }
    /* 4.4b freezing of rain, :1864-1908 */
    if (zqx[QR][k] > zepsec) {
      if (t <= rtt && ztp1[k - 1] > rtt) {
        zqpretot = fmax(zqx[QS][k] + zqx[QR][k], zepsec);
        rainfrac_toprfz = zqx[QR][k] / zqpretot;
      }
      if (t < rtt) {
        double zfrzmax;
        if (rainfrac_toprfz > 0.8) {
          const double zlambda = pow(y->rcl_fac1 / (zrho * zqx[QR][k]), y->rcl_fac2);
          const double ztemp = y->rcl_fzrab * (t - rtt);
          const double zfrz = ((ptsphy * (y->rcl_const5r / zrho)) * (exp(ztemp) - 1.0)) * pow(zlambda, y->rcl_const6r);
          zfrzmax = fmax(zfrz, 0.0);
        } else {
          const double zcons1 = fabs((ptsphy * (1.0 + 0.5 * (rtt - t))) / y->rtaumel);
          zfrzmax = fmax(((rtt - t) * zcons1) * zrldcp, 0.0);
        }
        if (zfrzmax > zepsec) {
          const double zfrz = fmin(zqx[QR][k], zfrzmax);
          zsolqa[QR][QS] += zfrz;
          zsolqa[QS][QR] -= zfrz;
        }
      }
    }
    /* 4.4c freezing of liquid, :1913-1928 */
    {
      const double zfrzmax = fmax((y->rthomo - t) * zrldcp, 0.0);
      const int m = QL, n = imelt[QL];
      if (zfrzmax > zepsec && zqxfg[m] > zepsec) {
        const double zfrz = fmin(zqxfg[m], zfrzmax);
        zsolqa[m][n] += zfrz;
        zsolqa[n][m] -= zfrz;
      }
    }
    /* 4.5 evaporation of rain, Abel & Boutle (ievaprain = 2), :1982-2040 */
    {
      double zzrh = y->rprecrhmax + ((1.0 - y->rprecrhmax) * zcovpmax) / fmax(zepsec, 1.0 - za[k]);
      zzrh = fmin(fmax(zzrh, y->rprecrhmax), 1.0);
      zzrh = fmin(0.8, zzrh);
      const double zqe = fmax(0.0, fmin(zqx[QV][k], zqsliq[k]));
      if (zcovpclr > zepsec && zqxfg[QR] > zepsec && zqe < zzrh * zqsliq[k]) {
        const double zpreclr = zqxfg[QR] / zcovptot;
        const double zfallcorr = pow(y->rdensref / zrho, 0.4);
        const double zesatliq = (rv / rd) * (r2es * f_eliq(t));
        const double zlambda = pow(y->rcl_fac1 / (zrho * zpreclr), y->rcl_fac2);
        const double zevap_denom = y->rcl_cdenom1 * zesatliq - y->rcl_cdenom2 * t * zesatliq + (y->rcl_cdenom3 * pow(t, 3.0)) * pap;
        const double zcorr2 = (pow(t / 273.0, 1.5) * 393.0) / (t + 120.0);
        const double zsubsat = fmax(zzrh * zqsliq[k] - zqe, 0.0);
        const double zbeta = (((((0.5 / zqsliq[k]) * sq2(t)) * zesatliq) * y->rcl_const1r) * (zcorr2 / zevap_denom)) *
                             (0.78 / pow(zlambda, y->rcl_const4r) +
                              (y->rcl_const2r * sqrt(zrho * zfallcorr)) / (sqrt(zcorr2) * pow(zlambda, y->rcl_const3r)));
        const double zdenom = 1.0 + zbeta * ptsphy;
        const double zdpevap = (((zcovpclr * zbeta) * ptsphy) * zsubsat) / zdenom;
        const double zevap = fmin(zdpevap, zqxfg[QR]);
        zsolqa[QR][QV] += zevap;
        zsolqa[QV][QR] -= zevap;
        zcovptot = fmax(y->rcovpmin, zcovptot - fmax(0.0, ((zcovptot - za[k]) * zevap) / zqxfg[QR]));
        zqxfg[QR] -= zevap;
      }
    }
    /* 4.5 evaporation of snow (ievapsnow = 1), :2048-2087 */
    {
      double zzrh = y->rprecrhmax + ((1.0 - y->rprecrhmax) * zcovpmax) / fmax(zepsec, 1.0 - za[k]);
      zzrh = fmin(fmax(zzrh, y->rprecrhmax), 1.0);
      double zqe = (zqx[QV][k] - za[k] * zqsice[k]) / fmax(zepsec, 1.0 - za[k]);
      zqe = fmax(0.0, fmin(zqe, zqsice[k]));
      if (zcovpclr > zepsec && zqxfg[QS] > zepsec && zqe < zzrh * zqsice[k]) {
        const double cd = zcovptot * zdtgdp;
        const double zpreclr = (zqxfg[QS] * zcovpclr) / copysign(fmax(fabs(cd), zepsilon), cd);
        const double zbeta1 = ((sqrt(pap / paph_surf) / y->rvrfactor) * zpreclr) / fmax(zcovpclr, zepsec);
        const double zbeta = (rg * y->rpecons) * pow(zbeta1, 0.5777);
        const double zdenom = 1.0 + (zbeta * ptsphy) * zcorqsice;
        const double zdpr = ((((zcovpclr * zbeta) * (zqsice[k] - zqe)) / zdenom) * zdp) * zrg_r;
        const double zdpevap = zdpr * zdtgdp;
        const double zevap = fmin(zdpevap, zqxfg[QS]);
        zsolqa[QS][QV] += zevap;
        zsolqa[QV][QS] -= zevap;
        zcovptot = fmax(y->rcovpmin, zcovptot - fmax(0.0, ((zcovptot - za[k]) * zevap) / zqxfg[QS]));
        zqxfg[QS] -= zevap;
      }
    }
    /* evaporate small precipitation amounts, :2144-2158 */
    for (int m = 0; m < NCLV; m++)
      if (llfall[m] && zqxfg[m] < y->rlmin) { zsolqa[m][QV] += zqxfg[m]; zsolqa[QV][m] -= zqxfg[m]; }

    /* 5.1 solver for cloud cover, :2168-2180 */
    double zanew = (za[k] + zsolac) / (1.0 + zsolab);
    zanew = fmin(zanew, 1.0);
    if (zanew < y->ramin) zanew = 0.0;
    const double zda = zanew - zaorig[k];
    zanewm1 = zanew;

    /* 5.2 truncate explicit sinks, species by species, :2233-2286 */
    for (int m = 0; m < NCLV; m++) {
      double psum = 0.0;
      for (int n = 0; n < NCLV; n++) psum += zsolqa[n][m];
      const double zsinksum = 0.0 - psum;
      const double zmm = fmax(zqx[m][k], zepsec);
      const double zrr = fmax(zsinksum, zmm);
      const double zratio = zmm / zrr;
      for (int n = 0; n < NCLV; n++)
        if (zsolqa[n][m] < 0.0) {
          zsolqa[n][m] = zsolqa[n][m] * zratio;
          zsolqa[m][n] = zsolqa[m][n] * zratio;
        }
    }
    /* 5.2.2 LHS, RHS, :2294-2339 */
    for (int m = 0; m < NCLV; m++)
      for (int n = 0; n < NCLV; n++) {
        if (n == m) {
          zqlhs[m][n] = 1.0 + zfallsink[m];
          for (int o = 0; o < NCLV; o++) zqlhs[m][n] += zsolqb[n][o];
        } else {
          zqlhs[m][n] = -zsolqb[m][n];
        }
      }
    for (int m = 0; m < NCLV; m++) {
      double zexplicit = 0.0;
      for (int n = 0; n < NCLV; n++) zexplicit += zsolqa[n][m];
      zqxn[m] = zqx[m][k] + zexplicit;
    }
    /* non-pivoting LU + substitutions, :2351-2397 */
    for (int n = 0; n < NCLV - 1; n++)
      for (int m = n + 1; m < NCLV; m++) {
        zqlhs[n][m] = zqlhs[n][m] / zqlhs[n][n];
        for (int i = n + 1; i < NCLV; i++) zqlhs[i][m] -= zqlhs[n][m] * zqlhs[i][n];
      }
    for (int n = 1; n < NCLV; n++)
      for (int m = 0; m < n; m++) zqxn[n] -= zqlhs[m][n] * zqxn[m];
    zqxn[NCLV - 1] = zqxn[NCLV - 1] / zqlhs[NCLV - 1][NCLV - 1];
    for (int n = NCLV - 2; n >= 0; n--) {
      for (int m = n + 1; m < NCLV; m++) zqxn[n] -= zqlhs[m][n] * zqxn[m];
      zqxn[n] = zqxn[n] / zqlhs[n][n];
    }
    /* clip, :2402-2412 */
    for (int n = 0; n < NCLV - 1; n++)
      if (zqxn[n] < zepsec) { zqxn[QV] += zqxn[n]; zqxn[n] = 0.0; }
    for (int m = 0; m < NCLV; m++) { zqxnm1[m] = zqxn[m]; zqxn2d[m][k] = zqxn[m]; }

    /* 5.3 fluxes to next level, :2430-2448 */
    for (int m = 0; m < NCLV; m++) zpfplsx[m][k + 1] = (zfallsink[m] * zqxn[m]) * zrdtgdp;
    zqpretot = zpfplsx[QS][k + 1] + zpfplsx[QR][k + 1];
    if (zqpretot < zepsec) zcovptot = 0.0;

    /* 6. tendencies, :2456-2506 */
    for (int m = 0; m < NCLV - 1; m++) {
      const double zfluxq = zpsupsatsrce[m] + zconvsrce[m] + zfallsrce[m] - (zfallsink[m] + zconvsink[m]) * zqxn[m];
      if (iphase[m] == 1) tlt[k] += (ralvdcp * (zqxn[m] - zqx[m][k] - zfluxq)) * zqtmst;
      if (iphase[m] == 2) tlt[k] += (ralsdcp * (zqxn[m] - zqx[m][k] - zfluxq)) * zqtmst;
      tlcld[m][k] += (zqxn[m] - zqx0[m][k]) * zqtmst;
    }
    tlq[k] += (zqxn[QV] - zqx[QV][k]) * zqtmst;
    tla[k] += zda * zqtmst;
    AT(c->pcovptot, k) = zcovptot;
  }

  /* ---- write tendencies (the reference accumulates them in place) --------------------------- */
  for (int k = 0; k < klev; k++) {
    AT(c->lt, k) = tlt[k];
    AT(c->lq, k) = tlq[k];
    AT(c->la, k) = tla[k];
    for (int m = 0; m < 4; m++) AT3(c->lcld, m, k) = tlcld[m][k];
  }
  *c->prainfrac = rainfrac_toprfz;

  /* ============================ 8. flux diagnostics, :2521-2582 ============================== */
  for (int k = 0; k <= klev; k++) {
    AT(c->fplsl, k) = zpfplsx[QR][k] + zpfplsx[QL][k];
    AT(c->fplsn, k) = zpfplsx[QS][k] + zpfplsx[QI][k];
  }
  AT(c->fsqlf, 0) = 0.0; AT(c->fsqif, 0) = 0.0; AT(c->fsqrf, 0) = 0.0; AT(c->fsqsf, 0) = 0.0;
  AT(c->fcqlng, 0) = 0.0; AT(c->fcqnng, 0) = 0.0; AT(c->fcqrng, 0) = 0.0; AT(c->fcqsng, 0) = 0.0;
  AT(c->fsqltur, 0) = 0.0; AT(c->fsqitur, 0) = 0.0;
  for (int k = 0; k < klev; k++) {
    const double zgdph_r = -zrg_r * (AT(c->paph, k + 1) - AT(c->paph, k)) * zqtmst;
    AT(c->fsqlf, k + 1) = AT(c->fsqlf, k);
    AT(c->fsqif, k + 1) = AT(c->fsqif, k);
    AT(c->fsqrf, k + 1) = AT(c->fsqlf, k);          /* sic: rain row copies the LIQUID sum, :2550 */
    AT(c->fsqsf, k + 1) = AT(c->fsqif, k);
    AT(c->fcqlng, k + 1) = AT(c->fcqlng, k);
    AT(c->fcqnng, k + 1) = AT(c->fcqnng, k);
    AT(c->fcqrng, k + 1) = AT(c->fcqlng, k);
    AT(c->fcqsng, k + 1) = AT(c->fcqnng, k);
    AT(c->fsqltur, k + 1) = AT(c->fsqltur, k);
    AT(c->fsqitur, k + 1) = AT(c->fsqitur, k);
    const double zalfaw = zfoealfa[k];
    AT(c->fsqlf, k + 1) += (zqxn2d[QL][k] - zqx0[QL][k] + AT(c->pvfl, k) * ptsphy - zalfaw * AT(c->plude, k)) * zgdph_r;
    AT(c->fcqlng, k + 1) += zlneg[QL][k] * zgdph_r;
    AT(c->fsqltur, k + 1) += (AT(c->pvfl, k) * ptsphy) * zgdph_r;
    AT(c->fsqrf, k + 1) += (zqxn2d[QR][k] - zqx0[QR][k]) * zgdph_r;
    AT(c->fcqrng, k + 1) += zlneg[QR][k] * zgdph_r;
    AT(c->fsqif, k + 1) += (zqxn2d[QI][k] - zqx0[QI][k] + AT(c->pvfi, k) * ptsphy - (1.0 - zalfaw) * AT(c->plude, k)) * zgdph_r;
    AT(c->fcqnng, k + 1) += zlneg[QI][k] * zgdph_r;
    AT(c->fsqitur, k + 1) += (AT(c->pvfi, k) * ptsphy) * zgdph_r;
    AT(c->fsqsf, k + 1) += (zqxn2d[QS][k] - zqx0[QS][k]) * zgdph_r;
    AT(c->fcqsng, k + 1) += zlneg[QS][k] * zgdph_r;
  }
  for (int k = 0; k <= klev; k++) {
    AT(c->fhpsl, k) = -rlvtt * AT(c->fplsl, k);
    AT(c->fhpsn, k) = -rlstt * AT(c->fplsn, k);
  }
#undef AT
#undef AT3
}

/* Same argument list as the reference kernel (cloudsc_c.h:18-29); returns 0, or -1 when the
 * parameter set / level count is outside the restated domain (the reference has no error path). */
int cloudsc_oracle(int kidia, int kfdia, int klon, int klev, double ptsphy,
  double *v_pt, double *v_pq,
  double *v_tendency_cml_t, double *v_tendency_cml_q, double *v_tendency_cml_a, double *v_tendency_cml_cld,
  double *v_tendency_tmp_t, double *v_tendency_tmp_q, double *v_tendency_tmp_a, double *v_tendency_tmp_cld,
  double *v_tendency_loc_t, double *v_tendency_loc_q, double *v_tendency_loc_a, double *v_tendency_loc_cld,
  double *v_pvfa, double *v_pvfl, double *v_pvfi, double *v_pdyna, double *v_pdynl, double *v_pdyni,
  double *v_phrsw, double *v_phrlw, double *v_pvervel, double *v_pap, double *v_paph, double *v_plsm,
  int *v_ktype, double *v_plu, double *v_plude, double *v_psnde, double *v_pmfu, double *v_pmfd,
  double *v_pa, double *v_pclv, double *v_psupsat, double *v_plcrit_aer, double *v_picrit_aer,
  double *v_pre_ice, double *v_pccn, double *v_pnice, double *v_pcovptot, double *v_prainfrac_toprfz,
  double *v_pfsqlf, double *v_pfsqif, double *v_pfcqnng, double *v_pfcqlng, double *v_pfsqrf,
  double *v_pfsqsf, double *v_pfcqrng, double *v_pfcqsng, double *v_pfsqltur, double *v_pfsqitur,
  double *v_pfplsl, double *v_pfplsn, double *v_pfhpsl, double *v_pfhpsn) {
  (void)v_tendency_cml_t; (void)v_tendency_cml_q; (void)v_tendency_cml_a; (void)v_tendency_cml_cld;
  (void)v_pvfa; (void)v_pdyna; (void)v_pdynl; (void)v_pdyni;
  (void)v_plcrit_aer; (void)v_pccn;
  if (!oracle_check_params() || klev > MAXLEV || klev < 2) return -1;
  for (int jl = kidia - 1; jl < kfdia; jl++) {
    struct colptr c;
    c.ls = (size_t)klon; c.ss = (size_t)klev * klon;
    c.pt = v_pt + jl; c.pq = v_pq + jl; c.tt = v_tendency_tmp_t + jl; c.tq = v_tendency_tmp_q + jl;
    c.ta = v_tendency_tmp_a + jl; c.tcld = v_tendency_tmp_cld + jl; c.pvfl = v_pvfl + jl; c.pvfi = v_pvfi + jl;
    c.phrsw = v_phrsw + jl; c.phrlw = v_phrlw + jl; c.pvervel = v_pvervel + jl; c.pap = v_pap + jl;
    c.paph = v_paph + jl; c.plu = v_plu + jl; c.psnde = v_psnde + jl; c.pmfu = v_pmfu + jl; c.pmfd = v_pmfd + jl;
    c.pa = v_pa + jl; c.pclv = v_pclv + jl; c.psupsat = v_psupsat + jl;
    c.pre_ice = v_pre_ice ? v_pre_ice + jl : NULL; c.picrit_aer = v_picrit_aer ? v_picrit_aer + jl : NULL;
    c.pnice = v_pnice ? v_pnice + jl : NULL;
    c.plsm = v_plsm[jl]; c.ktype = v_ktype[jl];
    c.plude = v_plude + jl; c.lt = v_tendency_loc_t + jl; c.lq = v_tendency_loc_q + jl; c.la = v_tendency_loc_a + jl;
    c.lcld = v_tendency_loc_cld + jl; c.pcovptot = v_pcovptot + jl; c.prainfrac = v_prainfrac_toprfz + jl;
    c.fsqlf = v_pfsqlf + jl; c.fsqif = v_pfsqif + jl; c.fcqnng = v_pfcqnng + jl; c.fcqlng = v_pfcqlng + jl;
    c.fsqrf = v_pfsqrf + jl; c.fsqsf = v_pfsqsf + jl; c.fcqrng = v_pfcqrng + jl; c.fcqsng = v_pfcqsng + jl;
    c.fsqltur = v_pfsqltur + jl; c.fsqitur = v_pfsqitur + jl; c.fplsl = v_pfplsl + jl; c.fplsn = v_pfplsn + jl;
    c.fhpsl = v_pfhpsl + jl; c.fhpsn = v_pfhpsn + jl;
    column(&c, klev, ptsphy);
  }
  return 0;
}
