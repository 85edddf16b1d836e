/* cloudsc_validate.h -- same entry point as src/cloudsc_c/cloudsc/cloudsc_validate.h:16-20 */
#ifndef CLOUDSC_VALIDATE_H
#define CLOUDSC_VALIDATE_H
#include "load_state.h"

void cloudsc_validate(const int nlon, const int nlev, const int nclv, const int ngptot, const int nproma,
                      double *plude, double *pcovptot, double *prainfrac_toprfz, double *pfsqlf, double *pfsqif,
                      double *pfcqlng, double *pfcqnng, double *pfsqrf, double *pfsqsf, double *pfcqrng, double *pfcqsng,
                      double *pfsqltur, double *pfsqitur, double *pfplsl, double *pfplsn, double *pfhpsl, double *pfhpsn,
                      double *tend_loc_a, double *tend_loc_q, double *tend_loc_t, double *tend_loc_cld);

/* Number of fields of the last cloudsc_validate() call that failed the parity bar of BASELINE.md
 * section 4 (sum|err|/sum|ref| <= 1e-10 and max|err| <= 1e-9 max|ref|); the reference only prints. */
int cloudsc_validate_failures(void);
#endif
