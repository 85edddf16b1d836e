/*
 * load_state.h -- same three entry points, argument for argument, as the reference loader
 * (src/cloudsc_c/cloudsc/load_state.h:22-39), so cloudsc_driver() reads like the reference's.
 * Files: "input.h5" / "reference.h5" in the working directory (the reference's INPUT_FILE /
 * REFERENCE_FILE, load_state.c:18-19), overridable with $CLOUDSC_INPUT_H5 / $CLOUDSC_REFERENCE_H5.
 * One deliberate difference: a missing or malformed file is a diagnosed failure (message on stderr,
 * exit status 2) instead of the reference's silent use of uninitialised memory.
 */
#ifndef CLOUDSC_LOAD_STATE_H
#define CLOUDSC_LOAD_STATE_H

#include <stdio.h>
#include <stdlib.h>
#include "cloudsc_params.h"

void query_state(int *klon, int *klev);

void load_state(const int nlon, const int nlev, const int nclv, const int ngptot, const int nproma,
                double *ptsphy, double *plcrit_aer, double *picrit_aer,
                double *pre_ice, double *pccn, double *pnice, double *pt, double *pq,
                double *tend_cml_t, double *tend_cml_q, double *tend_cml_a, double *tend_cml_cld,
                double *tend_tmp_t, double *tend_tmp_q, double *tend_tmp_a, double *tend_tmp_cld,
                double *pvfa, double *pvfl, double *pvfi, double *pdyna, double *pdynl, double *pdyni,
                double *phrsw, double *phrlw, double *pvervel, double *pap, double *paph, double *plsm,
                int *ktype, double *plu, double *plude, double *psnde, double *pmfu,
                double *pmfd, double *pa, double *pclv, double *psupsat);

void load_reference(const int nlon, const int nlev, const int nclv, const int ngptot, const int nproma,
                    double *plude, double *pcovptot, double *prainfrac_toprfz, double *pfsqlf, double *pfsqif,
                    double *pfcqlng, double *pfcqnng, double *pfsqrf, double *pfsqsf, double *pfcqrng, double *pfcqsng,
                    double *pfsqltur, double *pfsqitur, double *pfplsl, double *pfplsn, double *pfhpsl, double *pfhpsn,
                    double *tend_loc_a, double *tend_loc_q, double *tend_loc_t, double *tend_loc_cld);

/* extension used by the multi-GPU driver: global index of the first column this process holds
 * (column g of the expanded set is source column (first + g) mod nlon); default 0 */
void load_state_set_first_column(long long first);
#endif
