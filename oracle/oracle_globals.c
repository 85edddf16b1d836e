/* oracle_globals.c -- TEST INFRASTRUCTURE.  Parameter globals of the oracle port (the reference's
 * yoecldp_c.c / yomcst_c.c / yoethf_c.c play this role in oracle/_ref). */
#include "cloudsc_params.h"
int nclv, ncldql, ncldqi, ncldqr, ncldqs, ncldqv;
struct TECLDP *yrecldp;
#define CLOUDSC_X(n) double n;
CLOUDSC_YOMCST_REALS(CLOUDSC_X)
CLOUDSC_YOETHF_REALS(CLOUDSC_X)
#undef CLOUDSC_X
