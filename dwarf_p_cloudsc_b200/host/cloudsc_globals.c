/*
 * cloudsc_globals.c -- storage for the parameter globals of the dwarf-cloudsc-c contract
 * (reference: src/cloudsc_c/cloudsc/yoecldp_c.c, yomcst_c.c, yoethf_c.c), generated from the
 * X-macro tables of include/cloudsc_params.h.  load_state() fills them, cloudsc_c() /
 * cloudsc_b200_set_params_from_globals() snapshot them into __constant__ memory.
 */
#include "cloudsc_params.h"

int nclv, ncldql, ncldqi, ncldqr, ncldqs, ncldqv;
struct TECLDP *yrecldp;
#define CLOUDSC_X(n) double n;
CLOUDSC_YOMCST_REALS(CLOUDSC_X)
CLOUDSC_YOETHF_REALS(CLOUDSC_X)
#undef CLOUDSC_X
