/* mycpu.h -- core the calling thread runs on (reference: src/cloudsc_c/cloudsc/mycpu.h) */
#ifndef CLOUDSC_MYCPU_H
#define CLOUDSC_MYCPU_H
int mycpu(void);
int mycpu_(void);
#endif
