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
  X(rcl_const3r) X(rcl_const4r) X(rcl_fac1) X(rcl_fac2) X(rcl_const5r) X(rcl_const6r) X(rcl_fzrab)

template <typename T>
struct CloudscConst {
#define CLOUDSC_X(n) T n;
  CLOUDSC_DEV_CST(CLOUDSC_X)
  CLOUDSC_DEV_CLDP(CLOUDSC_X)
#undef CLOUDSC_X
  /* pow(rcl_kk_cloud_num_{sea,land}, rcl_kkbaun): per-run constants of cloudsc_c.c:1721 */
  T kk_num_pow_sea, kk_num_pow_land;
  int nssopt;
  int ncldtop;
};

#endif
