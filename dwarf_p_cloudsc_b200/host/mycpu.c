/* mycpu.c -- reference: src/cloudsc_c/cloudsc/mycpu.c:22-31 (sched_getcpu wrapper with a Fortran-style alias) */
#define _GNU_SOURCE
#include "mycpu.h"
#if defined(__linux__)
#include <sched.h>
int mycpu(void) { return sched_getcpu(); }
#else
int mycpu(void) { return 0; }
#endif
int mycpu_(void) { return mycpu(); }
