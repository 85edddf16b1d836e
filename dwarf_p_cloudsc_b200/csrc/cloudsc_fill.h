/*
 * cloudsc_fill.h -- host-side conversion  (struct TECLDP, struct cloudsc_consts) -> CloudscConst<T>,
 * plus the check that the parameter set is one the fused kernel is specialised for.
 * Shared by the C ABI (cloudsc_abi.cu) and the host-emulation test harness (tests/host_emul).
 */
#ifndef CLOUDSC_FILL_H
#define CLOUDSC_FILL_H

#include <cmath>
#include <cstdio>
#include "cloudsc_consts.h"
#include "cloudsc_params.h"

namespace cloudsc {

/* Returns NULL if supported, else a static message naming the offending member. */
static inline const char *check_params(const struct TECLDP *p) {
  /* laerliqautolsp / laerliqcoll only act inside the iwarmrain == 1 branch (cloudsc_c.c:1654,1679), which the
     reference hard-codes away (:367); the two live aerosol switches laericesed (:1561) and laericeauto (:1625) are
     implemented by the AER instantiation of the column code */
  if (p->laerliqautolsp) return "yrecldp->laerliqautolsp != 0 is not supported (iwarmrain == 1 only)";
  if (p->laerliqcoll) return "yrecldp->laerliqcoll != 0 is not supported (iwarmrain == 1 only)";
  if (p->nssopt < 0 || p->nssopt > 3) return "yrecldp->nssopt must be 0..3 (zqe is undefined otherwise, cloudsc_c.c:1275-1302)";
  if (p->ncldtop < 2) return "yrecldp->ncldtop must be >= 2 (level ncldtop reads level ncldtop-1, cloudsc_c.c:1390,1597,1868)";
  /* llfall pattern the unrolled species code assumes: rain and snow fall, liquid does not;
     ice sediments but is flagged non-falling (cloudsc_c.c:437-456) */
  if (!(p->rvrain > 0.0)) return "yrecldp->rvrain must be > 0";
  if (!(p->rvsnow > 0.0)) return "yrecldp->rvsnow must be > 0";
  return nullptr;
}

template <typename T>
static inline void fill_consts(const struct TECLDP *p, const struct cloudsc_consts *k, CloudscConst<T> *out) {
#define CLOUDSC_X(n) out->n = (T)k->n;
  CLOUDSC_DEV_CST(CLOUDSC_X)
#undef CLOUDSC_X
#define CLOUDSC_X(n) out->n = (T)p->n;
  CLOUDSC_DEV_CLDP(CLOUDSC_X)
#undef CLOUDSC_X
#define CLOUDSC_X(n, v) out->n = (T)(v);
  CLOUDSC_DEV_LITERALS(CLOUDSC_X)
#undef CLOUDSC_X
  out->zepsilon = (T)100.0 * (sizeof(T) == 8 ? (T)2.220446049250313e-16 : (T)1.1920929e-07f);
  out->r_rd = (T)(1.0 / k->rd);
  out->r_rv = (T)(1.0 / k->rv);
  out->r_rtaumel = (T)(1.0 / p->rtaumel);
  out->r_rlcritsnow = (T)(1.0 / p->rlcritsnow);
  out->r_rdepliqrefdepth = (T)(1.0 / p->rdepliqrefdepth);
  out->r_rvrfactor = (T)(1.0 / p->rvrfactor);
  out->r_15000 = (T)(1.0 / 15000.0);
  out->r_273 = (T)(1.0 / 273.0);
  out->r_p02 = (T)(1.0 / 0.2);
  out->r_p0024 = (T)(1.0 / 0.024);
  out->kk_num_pow_sea = (T)std::pow(p->rcl_kk_cloud_num_sea, p->rcl_kkbaun);
  out->kk_num_pow_land = (T)std::pow(p->rcl_kk_cloud_num_land, p->rcl_kkbaun);
  out->nssopt = p->nssopt;
  out->ncldtop = p->ncldtop;
  out->laericesed = p->laericesed ? 1 : 0;
  out->laericeauto = p->laericeauto ? 1 : 0;
}

/* bit 0: PRE_ICE is read (laericesed), bit 1: PICRIT_AER and PNICE are read (laericeauto) */
static inline int aerosol_mask(const struct TECLDP *p) { return (p->laericesed ? 1 : 0) | (p->laericeauto ? 2 : 0); }

/* Per-launch constants; every member is computed in T with the reference's own expression. */
template <typename T>
static inline void fill_step(const CloudscConst<T> &c, double ptsphy_d, StepConst<T> *s) {
  const T one = (T)1.0, ptsphy = (T)ptsphy_d;
  s->ptsphy = ptsphy;
  s->zqtmst = one / ptsphy;
  s->zrdcp = c.rd / c.rcpd;
  s->zrdcp_half = s->zrdcp * (T)0.5;
  s->zrg_r = one / c.rg;
  s->neg_zrg_r = -s->zrg_r;
  s->zrldcp = one / (c.ralsdcp - c.ralvdcp);
  s->r_ptsphy_rg = one / (ptsphy * c.rg);
  s->half_rg = (T)0.5 * c.rg;
  s->ldifdt = c.rcldiff * ptsphy;
  s->ldifdt_conv = c.rcldiff_convi * s->ldifdt;
  s->faci_koop = ptsphy / c.rkooptau;
  s->snowlin = ptsphy * c.rsnowlin1;
  s->rv_rd = c.rv / c.rd;
  s->rg_rpecons = c.rg * c.rpecons;
  s->one_m_ramin = one - c.ramin;
  s->one_m_zepsec = one - c.zepsec;
  s->one_m_ramid = one - c.ramid;
  s->one_m_rprecrhmax = one - c.rprecrhmax;
}

} /* namespace cloudsc */
#endif
