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
  /* branches the reference only takes when these flags are set read extra fields and, for
     laericesed, leak a per-column value across columns (cloudsc_c.c:1561-1566) */
  if (p->laerliqautolsp) return "yrecldp->laerliqautolsp != 0 is not supported";
  if (p->laerliqcoll) return "yrecldp->laerliqcoll != 0 is not supported";
  if (p->laericesed) return "yrecldp->laericesed != 0 is not supported";
  if (p->laericeauto) return "yrecldp->laericeauto != 0 is not supported";
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
  out->kk_num_pow_sea = (T)std::pow(p->rcl_kk_cloud_num_sea, p->rcl_kkbaun);
  out->kk_num_pow_land = (T)std::pow(p->rcl_kk_cloud_num_land, p->rcl_kkbaun);
  out->nssopt = p->nssopt;
  out->ncldtop = p->ncldtop;
}

} /* namespace cloudsc */
#endif
