/*
 * cloudsc_params.h -- parameter objects of the CLOUDSC dwarf, ABI-compatible with the reference.
 *
 * Mirrors, field for field and in the same order, the three parameter "modules" of
 * dwarf-cloudsc-c so that objects filled by either side can be handed to the other:
 *
 *   struct TECLDP / yrecldp   <- src/cloudsc_c/cloudsc/yoecldp_c.h:21-141   (912 bytes on LP64)
 *   YOMCST globals (rg, rd..) <- src/cloudsc_c/cloudsc/yomcst_c.h:14-72
 *   YOETHF globals (r2es, ..) <- src/cloudsc_c/cloudsc/yoethf_c.h:14-36
 *
 * The member lists are X-macros so that the loader (host/load_state.c), the constant-memory
 * snapshot (csrc/cloudsc_abi.cu) and the ctypes mirror (dwarf_p_cloudsc_b200/params.py, checked
 * against sizeof/offsetof in tests) are generated from ONE table.
 */
#ifndef CLOUDSC_PARAMS_H
#define CLOUDSC_PARAMS_H

#ifdef __cplusplus
extern "C" {
#endif

/* -- YRECLDP: 102 leading doubles, in declaration order --------------------------------- */
#define CLOUDSC_TECLDP_REALS(X) \
  X(ramid) X(rcldiff) X(rcldiff_convi) X(rclcrit) X(rclcrit_sea) X(rclcrit_land) X(rkconv) \
  X(rprc1) X(rprc2) X(rcldmax) X(rpecons) X(rvrfactor) X(rprecrhmax) X(rtaumel) X(ramin) \
  X(rlmin) X(rkooptau) X(rcldtopp) X(rlcritsnow) X(rsnowlin1) X(rsnowlin2) X(ricehi1) \
  X(ricehi2) X(riceinit) X(rvice) X(rvrain) X(rvsnow) X(rthomo) X(rcovpmin) X(rccn) X(rnice) \
  X(rccnom) X(rccnss) X(rccnsu) X(rcldtopcf) X(rdepliqrefrate) X(rdepliqrefdepth) \
  X(rcl_kkaac) X(rcl_kkbac) X(rcl_kkaau) X(rcl_kkbauq) X(rcl_kkbaun) \
  X(rcl_kk_cloud_num_sea) X(rcl_kk_cloud_num_land) \
  X(rcl_ai) X(rcl_bi) X(rcl_ci) X(rcl_di) X(rcl_x1i) X(rcl_x2i) X(rcl_x3i) X(rcl_x4i) \
  X(rcl_const1i) X(rcl_const2i) X(rcl_const3i) X(rcl_const4i) X(rcl_const5i) X(rcl_const6i) \
  X(rcl_apb1) X(rcl_apb2) X(rcl_apb3) \
  X(rcl_as) X(rcl_bs) X(rcl_cs) X(rcl_ds) X(rcl_x1s) X(rcl_x2s) X(rcl_x3s) X(rcl_x4s) \
  X(rcl_const1s) X(rcl_const2s) X(rcl_const3s) X(rcl_const4s) X(rcl_const5s) X(rcl_const6s) \
  X(rcl_const7s) X(rcl_const8s) \
  X(rdenswat) X(rdensref) X(rcl_ar) X(rcl_br) X(rcl_cr) X(rcl_dr) X(rcl_x1r) X(rcl_x2r) \
  X(rcl_x4r) X(rcl_ka273) X(rcl_cdenom1) X(rcl_cdenom2) X(rcl_cdenom3) X(rcl_schmidt) \
  X(rcl_dynvisc) X(rcl_const1r) X(rcl_const2r) X(rcl_const3r) X(rcl_const4r) X(rcl_fac1) \
  X(rcl_fac2) X(rcl_const5r) X(rcl_const6r) X(rcl_fzrab) X(rcl_fzrbb)

/* -- YRECLDP: the 17 ints that follow (logicals are stored as int) ------------------------ */
#define CLOUDSC_TECLDP_INTS(X) \
  X(lcldextra) X(lcldbudget) X(nssopt) X(ncldtop) X(naeclbc) X(naecldu) X(naeclom) X(naeclss) \
  X(naeclsu) X(nclddiag) X(naercld) X(laerliqautolsp) X(laerliqautocp) X(laerliqautocpb) \
  X(laerliqcoll) X(laericesed) X(laericeauto)

struct TECLDP {
#define CLOUDSC_X(n) double n;
  CLOUDSC_TECLDP_REALS(CLOUDSC_X)
#undef CLOUDSC_X
#define CLOUDSC_X(n) int n;
  CLOUDSC_TECLDP_INTS(CLOUDSC_X)
#undef CLOUDSC_X
  double nshapep;
  double nshapeq;
  int nbeta;
  /* the reference closes the struct with two zero-length [0][100] arrays (rbeta, rbetap1);
     they occupy no storage, so sizeof == 912 either way */
};

/* -- YOMCST: every global the reference header declares (order as declared) --------------- */
#define CLOUDSC_YOMCST_REALS(X) \
  X(rpi) X(rclum) X(rhpla) X(rkbol) X(rnavo) X(rday) X(rdayi) X(rhour) X(rea) X(repsm) \
  X(rsiyea) X(rsiday) X(romega) X(ra) X(rg) X(r1sa) X(rsigma) X(ri0) X(r) X(rmd) X(rmv) \
  X(rmo3) X(rd) X(rv) X(rcpd) X(rcpv) X(rcvd) X(rcvv) X(rkappa) X(retv) X(rmco2) X(rmch4) \
  X(rmn2o) X(rmco) X(rmhcho) X(rmno2) X(rmso2) X(rmso4) X(rcw) X(rcs) X(ratm) X(rtt) \
  X(rlvtt) X(rlstt) X(rlvzer) X(rlszer) X(rlmlt) X(rdt) X(restt) X(rgamw) X(rbetw) X(ralpw) \
  X(rgams) X(rbets) X(ralps) X(ralpd) X(rbetd) X(rgamd) X(rsnan)

/* -- YOETHF ------------------------------------------------------------------------------- */
#define CLOUDSC_YOETHF_REALS(X) \
  X(r2es) X(r3les) X(r3ies) X(r4les) X(r4ies) X(r5les) X(r5ies) X(rvtmp2) X(rhoh2o) \
  X(r5alvcp) X(r5alscp) X(ralvdcp) X(ralsdcp) X(ralfdcp) X(rtwat) X(rtber) X(rtbercu) \
  X(rtice) X(rticecu) X(rtwat_rtice_r) X(rtwat_rticecu_r) X(rkoop1) X(rkoop2)

/* The subset of YOMCST/YOETHF the kernel actually reads (cloudsc_c.c), in the order
 * load_state.c:540-567 fills them.  This is the `consts` block of the thin GPU ABI. */
#define CLOUDSC_KERNEL_CONSTS(X) \
  X(rg) X(rd) X(rcpd) X(retv) X(rlvtt) X(rlstt) X(rlmlt) X(rtt) X(rv) \
  X(r2es) X(r3les) X(r3ies) X(r4les) X(r4ies) X(r5les) X(r5ies) X(r5alvcp) X(r5alscp) \
  X(ralvdcp) X(ralsdcp) X(ralfdcp) X(rtwat) X(rtice) X(rticecu) X(rtwat_rtice_r) \
  X(rtwat_rticecu_r) X(rkoop1) X(rkoop2)

struct cloudsc_consts {
#define CLOUDSC_X(n) double n;
  CLOUDSC_KERNEL_CONSTS(CLOUDSC_X)
#undef CLOUDSC_X
};

/* Globals, exactly as the reference exposes them (yoecldp_c.h:14-19,143; yomcst_c.h; yoethf_c.h).
 * Defined in host/cloudsc_globals.c. */
extern int nclv, ncldql, ncldqi, ncldqr, ncldqs, ncldqv;
extern struct TECLDP *yrecldp;
#define CLOUDSC_X(n) extern double n;
CLOUDSC_YOMCST_REALS(CLOUDSC_X)
CLOUDSC_YOETHF_REALS(CLOUDSC_X)
#undef CLOUDSC_X

#ifdef __cplusplus
}
#endif
#endif /* CLOUDSC_PARAMS_H */
