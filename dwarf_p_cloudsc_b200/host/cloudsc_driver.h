/* cloudsc_driver.h -- same entry point as src/cloudsc_c/cloudsc/cloudsc_driver.h:27 */
#ifndef CLOUDSC_DRIVER_H
#define CLOUDSC_DRIVER_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cloudsc_b200.h"
#include "cloudsc_validate.h"
#include "load_state.h"

/* numthreads = number of GPUs to drive (one OpenMP host thread per GPU; <= 0 or more than are
 * visible -> all visible devices); numcols = NGPTOT; nproma = NPROMA.  Returns nothing, like the
 * reference; cloudsc_driver_status() is non-zero after a failed run or failed validation. */
void cloudsc_driver(int numthreads, int numcols, int nproma);
int cloudsc_driver_status(void);
#endif
