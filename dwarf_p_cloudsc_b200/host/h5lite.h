/*
 * h5lite.h -- dependency-free reader for the HDF5 subset the CLOUDSC dwarf's input.h5 / reference.h5
 * use (flat root group of contiguous, unfiltered, little-endian float64 / int32 / int64 / uint8
 * datasets; superblock v0, v1 object headers, group B-tree v1 + SNOD + local heap, layout v3).
 * Stands in for the libhdf5 calls of the reference loader (src/cloudsc_c/cloudsc/load_state.c:26-41:
 * H5Fopen / H5Dopen2 / H5Dread(H5T_NATIVE_DOUBLE | H5T_NATIVE_INT) / H5Fclose); libhdf5 does not
 * exist in this image.  C twin of dwarf_p_cloudsc_b200/h5lite.py (tests compare the two).
 */
#ifndef CLOUDSC_H5LITE_H
#define CLOUDSC_H5LITE_H
#include <stddef.h>

typedef struct h5lite_file h5lite_file;

/* element type classes of a dataset */
enum { H5LITE_F64 = 1, H5LITE_F32, H5LITE_I32, H5LITE_I64, H5LITE_U8, H5LITE_I8, H5LITE_I16, H5LITE_U16, H5LITE_U32, H5LITE_U64 };

typedef struct {
  int type;            /* H5LITE_* */
  int rank;
  size_t dims[8];
  size_t count;        /* product of dims (1 for rank 0) */
  const void *data;    /* points into the mapped file image; NULL for a never-written dataset (= zeros) */
} h5lite_dataset;

/* returns NULL on failure; h5lite_error() describes why */
h5lite_file *h5lite_open(const char *path);
void h5lite_close(h5lite_file *f);
const char *h5lite_error(void);
/* 0 on success, -1 if the dataset does not exist or is not a simple dataset */
int h5lite_find(const h5lite_file *f, const char *name, h5lite_dataset *out);
/* read `count` elements converted to double / int (the H5Dread(H5T_NATIVE_DOUBLE|INT) of the reference);
 * 0 on success, -1 on error (missing dataset, element count mismatch) */
int h5lite_read_double(const h5lite_file *f, const char *name, double *out, size_t count);
int h5lite_read_int(const h5lite_file *f, const char *name, int *out, size_t count);
/* number of root-level objects and the name of the i-th (B-tree order) */
size_t h5lite_count(const h5lite_file *f);
const char *h5lite_name(const h5lite_file *f, size_t i);
#endif
