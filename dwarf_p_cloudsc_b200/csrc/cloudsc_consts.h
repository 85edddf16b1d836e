/*
 * cloudsc_consts.h -- the parameter block the device code reads (lives in __constant__ memory).
 *
 * Built on the host from `struct TECLDP` + `struct cloudsc_consts` (include/cloudsc_params.h) by
 * `fill_device_consts()` in cloudsc_abi.cu; holds exactly the members cloudsc_c.c dereferences
 *  converted to the arithmetic type T, so the kernel never touches global memory
 * for a parameter (the reference CUDA variant dereferences a global `yrecldp` pointer at every use,
 * src/cloudsc_cuda/cloudsc/cloudsc_c_k_caching.cu:35,365,517).
 */
#ifndef CLOUDSC_CONSTS_H
#define CLOUDSC_CONSTS_H

/* members copied 1:1 from YOMCST/YOETHF */
#define CLOUDSC_DEV_CST(X) \
  X(rg) X(rd) X(rcpd) X(retv) X(rlvtt) X(rlstt) X(rlmlt) X(rtt) X(rv) \
  X(r2es) X(r3les) X(r3ies) X(r4les) X(r4ies) X(r5les) X(r5ies) X(r5alvcp) X(r5alscp) \
  X(ralvdcp) X(ralsdcp) X(ralfdcp) X(rtwat) X(rtice) X(rtwat_rtice_r) X(rkoop1) X(rkoop2)

/* members copied 1:1 from YRECLDP (only those the shipped code path reads) */
#define CLOUDSC_DEV_CLDP(X) \
  X(ramid) X(rcldiff) X(rcldiff_convi) X(rclcrit_sea) X(rclcrit_land) X(rpecons) X(rvrfactor) \
  X(rprecrhmax) X(rtaumel) X(ramin) X(rlmin) X(rkooptau) X(rlcritsnow) X(rsnowlin1) X(rsnowlin2) \
  X(riceinit) X(rvice) X(rvrain) X(rvsnow) X(rthomo) X(rcovpmin) X(rcldtopcf) X(rdepliqrefrate) \
  X(rdepliqrefdepth) X(rcl_kkaac) X(rcl_kkbac) X(rcl_kkaau) X(rcl_kkbauq) X(rcl_kkbaun) \
  X(rcl_kk_cloud_num_sea) X(rcl_kk_cloud_num_land) X(rcl_const1s) X(rcl_const7s) X(rcl_const8s) \
  X(rdensref) X(rcl_cdenom1) X(rcl_cdenom2) X(rcl_cdenom3) X(rcl_const1r) X(rcl_const2r) \
  X(rcl_const3r) X(rcl_const4r) X(rcl_fac1) X(rcl_fac2) X(rcl_const5r) X(rcl_const6r) X(rcl_fzrab) \
  X(rnice)

/*
 * Literals of cloudsc_c.c whose binary64 pattern has a non-zero low word.  FP64 instructions can
 * only encode the high word of an immediate, so such literals would each cost two move
 * instructions at every use; as members of the constant block they are free `c[bank][offset]`
 * operands.  (name, value)
 */
#define CLOUDSC_DEV_LITERALS(X) \
  X(zepsec, 1.0e-14) X(one_m_1em6, 1.0 - 1.0e-6) X(p0666, 0.666) X(p1296, 12.96) X(p0639, 0.639) \
  X(p0024, 0.024) X(p221, 2.21) X(p78, 7.8) X(p887, 8.87) X(p08, 0.8) X(p02, 0.2) X(p099, 0.99) \
  X(p001, 0.01) X(p03, 0.3) X(p04, 0.4) X(p05777, 0.5777) X(p078, 0.78) X(ztw1, 1329.31) \
  X(ztw2, 0.0074615) X(ztw4, 40.637) X(p0002, 0.002) X(p0333, 0.333)

template <typename T>
struct CloudscConst {
#define CLOUDSC_X(n) T n;
  CLOUDSC_DEV_CST(CLOUDSC_X)
  CLOUDSC_DEV_CLDP(CLOUDSC_X)
#undef CLOUDSC_X
#define CLOUDSC_X(n, v) T n;
  CLOUDSC_DEV_LITERALS(CLOUDSC_X)
#undef CLOUDSC_X
  T zepsilon;                           /* 100*epsilon(T), cloudsc_c.c:360 */
  /* pow(rcl_kk_cloud_num_{sea,land}, rcl_kkbaun): per-run constants of cloudsc_c.c:1721 */
  T kk_num_pow_sea, kk_num_pow_land;
  /* reciprocals of run constants (used by the fast-arithmetic build only) */
  T r_rd, r_rv, r_rtaumel, r_rlcritsnow, r_rdepliqrefdepth, r_rvrfactor, r_15000, r_273, r_p02, r_p0024;
  int nssopt;
  int ncldtop;
  /* aerosol switches the AER instantiation of the column code honours (cloudsc_c.c:1561,1625) */
  int laericesed, laericeauto;
};

/*
 * Per-launch constants: everything cloudsc_c.c derives from PTSPHY and the parameter objects before
 * its loops (:385-391 and the loop-invariant sub-expressions it re-evaluates per point).  Passed to
 * the kernel BY VALUE, i.e. it lives in the kernel-parameter constant bank.  Each member is exactly
 * the sub-expression the reference evaluates (same operands, same association), so precomputing it
 * on the host does not change a single bit of the result.
 */
template <typename T>
struct StepConst {
  T ptsphy;
  T zqtmst;        /* 1/ptsphy                    :385 */
  T zrdcp;         /* rd/rcpd                     :387 */
  T zrdcp_half;    /* zrdcp*0.5                   :1025 */
  T zrg_r;         /* 1/rg                        :390 */
  T neg_zrg_r;     /* -zrg_r                      :2547 */
  T zrldcp;        /* 1/(ralsdcp-ralvdcp)         :391 */
  T r_ptsphy_rg;   /* 1/(ptsphy*rg)               :807 */
  T half_rg;       /* 0.5*rg                      :1146 */
  T ldifdt;        /* rcldiff*ptsphy              :1088 */
  T ldifdt_conv;   /* rcldiff_convi*ldifdt        :1091 */
  T faci_koop;     /* ptsphy/rkooptau             :889 */
  T snowlin;       /* ptsphy*rsnowlin1            :1624 */
  T rv_rd;         /* rv/rd                       :2011 */
  T rg_rpecons;    /* rg*rpecons                  :2065 */
  T one_m_ramin;   /* 1-ramin                     :898 */
  T one_m_zepsec;  /* 1-zepsec                    :1254 */
  T one_m_ramid;   /* 1-ramid                     :1263 */
  T one_m_rprecrhmax; /* 1-rprecrhmax             :1990 */
};

#endif
