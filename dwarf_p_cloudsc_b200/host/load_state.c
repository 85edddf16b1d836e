/*
 * load_state.c -- query_state / load_state / load_reference of the dwarf-cloudsc-c driver contract.
 *
 * Behaviour follows src/cloudsc_c/cloudsc/load_state.c of the reference:
 *   - the input set holds NLON (=100) columns; NGPTOT columns are produced by replication, destination
 *     column g = b*nproma + i  <-  source column g mod nlon (:69-184); the padding columns of a partly
 *     filled last block are filled the same way (the zeroing at :84-92 is commented out upstream);
 *   - fields are stored [block][level][nproma] (2-D), [block][species][level][nproma] (3-D),
 *     [block][nproma] (1-D), which is how the file's C-order (level, column) arrays are laid out too;
 *   - every scalar is its own shape-(1,) dataset; ints and logicals are read as native int (:27-33);
 *   - the YOMCST / YOETHF globals and *yrecldp are filled from the file (:540-692).
 * The HDF5 back end is host/h5lite.c (no libhdf5 in this image).
 */
#include "load_state.h"

#include <ctype.h>
#include <string.h>

#include "h5lite.h"

static long long g_first_column = 0;
void load_state_set_first_column(long long first) { g_first_column = first; }

static const char *input_path(void) { const char *p = getenv("CLOUDSC_INPUT_H5"); return p && *p ? p : "input.h5"; }
static const char *reference_path(void) { const char *p = getenv("CLOUDSC_REFERENCE_H5"); return p && *p ? p : "reference.h5"; }

static void die(const char *what, const char *path) {
  fprintf(stderr, "load_state: %s (%s): %s\n", what, path, h5lite_error());
  exit(2);
}
static h5lite_file *open_or_die(const char *path) {
  h5lite_file *f = h5lite_open(path);
  if (!f) die("cannot open", path);
  return f;
}

void query_state(int *klon, int *klev) {
  h5lite_file *f = open_or_die(input_path());
  if (h5lite_read_int(f, "/KLEV", klev, 1) || h5lite_read_int(f, "/KLON", klon, 1)) die("KLON/KLEV", input_path());
  h5lite_close(f);
}

/* replicate `rows` rows of nlon source columns into [nblocks][rows][nproma] */
static void expand_rows(const double *src, double *dst, int rows, int nlon, int nproma, int nblocks) {
#pragma omp parallel for schedule(static)
  for (int b = 0; b < nblocks; ++b)
    for (int r = 0; r < rows; ++r) {
      double *d = dst + ((size_t)b * rows + r) * nproma;
      const double *s = src + (size_t)r * nlon;
      long long g = (g_first_column + (long long)b * nproma) % nlon;
      for (int i = 0; i < nproma; ++i) {
        d[i] = s[g];
        if (++g == nlon) g = 0;
      }
    }
}
static void expand_rows_int(const int *src, int *dst, int nlon, int nproma, int nblocks) {
  for (int b = 0; b < nblocks; ++b) {
    long long g = (g_first_column + (long long)b * nproma) % nlon;
    for (int i = 0; i < nproma; ++i) {
      dst[(size_t)b * nproma + i] = src[g];
      if (++g == nlon) g = 0;
    }
  }
}

static void load_and_expand(h5lite_file *f, const char *path, const char *name, int rows, int nlon, int nproma,
                            int nblocks, double *field, double *scratch) {
  if (!field) return;   /* callers may pass NULL for fields the kernel never touches */
  if (h5lite_read_double(f, name, scratch, (size_t)rows * nlon)) die(name, path);
  expand_rows(scratch, field, rows, nlon, nproma, nblocks);
}

static void upper_name(char *dst, size_t n, const char *prefix, const char *member) {
  /* the seven Khairoutdinov-Kogan members keep their mixed case in the file (serialbox names,
   * load_state.c:597-603) */
  static const char *const mixed[][2] = {
      {"rcl_kkaac", "RCL_KKAac"}, {"rcl_kkbac", "RCL_KKBac"}, {"rcl_kkaau", "RCL_KKAau"}, {"rcl_kkbauq", "RCL_KKBauq"},
      {"rcl_kkbaun", "RCL_KKBaun"}, {"rcl_kk_cloud_num_sea", "RCL_KK_cloud_num_sea"},
      {"rcl_kk_cloud_num_land", "RCL_KK_cloud_num_land"}};
  for (size_t i = 0; i < sizeof mixed / sizeof mixed[0]; ++i)
    if (!strcmp(member, mixed[i][0])) { snprintf(dst, n, "/%s%s", prefix, mixed[i][1]); return; }
  size_t k = (size_t)snprintf(dst, n, "/%s", prefix);
  for (; *member && k + 1 < n; ++member) dst[k++] = (char)toupper((unsigned char)*member);
  dst[k] = 0;
}

void load_state(const int nlon, const int nlev, const int nclv, const int ngptot, const int nproma,
                double *ptsphy, double *plcrit_aer, double *picrit_aer,
                double *pre_ice, double *pccn, double *pnice, double *pt, double *pq,
                double *tend_cml_t, double *tend_cml_q, double *tend_cml_a, double *tend_cml_cld,
                double *tend_tmp_t, double *tend_tmp_q, double *tend_tmp_a, double *tend_tmp_cld,
                double *pvfa, double *pvfl, double *pvfi, double *pdyna, double *pdynl, double *pdyni,
                double *phrsw, double *phrlw, double *pvervel, double *pap, double *paph, double *plsm,
                int *ktype, double *plu, double *plude, double *psnde, double *pmfu,
                double *pmfd, double *pa, double *pclv, double *psupsat) {
  const char *path = input_path();
  h5lite_file *f = open_or_die(path);
  const int nblocks = ngptot / nproma + (ngptot % nproma ? 1 : 0);
  double *scratch = malloc(sizeof(double) * (size_t)nclv * (nlev + 1) * nlon);
  int *iscratch = malloc(sizeof(int) * (size_t)nlon);
  if (!scratch || !iscratch) { fprintf(stderr, "load_state: out of memory\n"); exit(2); }

#define L2D(name, field) load_and_expand(f, path, name, nlev, nlon, nproma, nblocks, field, scratch)
#define L3D(name, field) load_and_expand(f, path, name, nclv * nlev, nlon, nproma, nblocks, field, scratch)
  L2D("PLCRIT_AER", plcrit_aer); L2D("PICRIT_AER", picrit_aer); L2D("PRE_ICE", pre_ice); L2D("PCCN", pccn);
  L2D("PNICE", pnice); L2D("PT", pt); L2D("PQ", pq);
  L2D("TENDENCY_CML_T", tend_cml_t); L2D("TENDENCY_CML_Q", tend_cml_q); L2D("TENDENCY_CML_A", tend_cml_a);
  L3D("TENDENCY_CML_CLD", tend_cml_cld);
  L2D("TENDENCY_TMP_T", tend_tmp_t); L2D("TENDENCY_TMP_Q", tend_tmp_q); L2D("TENDENCY_TMP_A", tend_tmp_a);
  L3D("TENDENCY_TMP_CLD", tend_tmp_cld);
  L2D("PVFA", pvfa); L2D("PVFL", pvfl); L2D("PVFI", pvfi); L2D("PDYNA", pdyna); L2D("PDYNL", pdynl); L2D("PDYNI", pdyni);
  L2D("PHRSW", phrsw); L2D("PHRLW", phrlw); L2D("PVERVEL", pvervel); L2D("PAP", pap);
  load_and_expand(f, path, "PAPH", nlev + 1, nlon, nproma, nblocks, paph, scratch);
  load_and_expand(f, path, "PLSM", 1, nlon, nproma, nblocks, plsm, scratch);
  if (ktype) {
    if (h5lite_read_int(f, "KTYPE", iscratch, (size_t)nlon)) die("KTYPE", path);
    expand_rows_int(iscratch, ktype, nlon, nproma, nblocks);
  }
  L2D("PLU", plu); L2D("PLUDE", plude); L2D("PSNDE", psnde); L2D("PMFU", pmfu); L2D("PMFD", pmfd); L2D("PA", pa);
  L3D("PCLV", pclv);
  L2D("PSUPSAT", psupsat);
#undef L2D
#undef L3D

  if (h5lite_read_double(f, "/PTSPHY", ptsphy, 1)) die("PTSPHY", path);
  char name[96];
  /* YOMCST / YOETHF members the kernel reads (load_state.c:540-567) */
#define CLOUDSC_X(n) upper_name(name, sizeof name, "", #n); if (h5lite_read_double(f, name, &n, 1)) die(name, path);
  CLOUDSC_KERNEL_CONSTS(CLOUDSC_X)
#undef CLOUDSC_X
  /* *yrecldp (load_state.c:569-690) */
  if (!yrecldp) { fprintf(stderr, "load_state: yrecldp is NULL (allocate it before load_state)\n"); exit(2); }
#define CLOUDSC_X(n) upper_name(name, sizeof name, "YRECLDP_", #n); if (h5lite_read_double(f, name, &yrecldp->n, 1)) die(name, path);
  CLOUDSC_TECLDP_REALS(CLOUDSC_X)
#undef CLOUDSC_X
#define CLOUDSC_X(n) upper_name(name, sizeof name, "YRECLDP_", #n); if (h5lite_read_int(f, name, &yrecldp->n, 1)) die(name, path);
  CLOUDSC_TECLDP_INTS(CLOUDSC_X)
#undef CLOUDSC_X
  if (h5lite_read_double(f, "/YRECLDP_NSHAPEP", &yrecldp->nshapep, 1)) die("YRECLDP_NSHAPEP", path);
  if (h5lite_read_double(f, "/YRECLDP_NSHAPEQ", &yrecldp->nshapeq, 1)) die("YRECLDP_NSHAPEQ", path);
  if (h5lite_read_int(f, "/YRECLDP_NBETA", &yrecldp->nbeta, 1)) die("YRECLDP_NBETA", path);

  free(scratch);
  free(iscratch);
  h5lite_close(f);
}

void load_reference(const int nlon, const int nlev, const int nclv, const int ngptot, const int nproma,
                    double *plude, double *pcovptot, double *prainfrac_toprfz, double *pfsqlf, double *pfsqif,
                    double *pfcqlng, double *pfcqnng, double *pfsqrf, double *pfsqsf, double *pfcqrng, double *pfcqsng,
                    double *pfsqltur, double *pfsqitur, double *pfplsl, double *pfplsn, double *pfhpsl, double *pfhpsn,
                    double *tend_loc_a, double *tend_loc_q, double *tend_loc_t, double *tend_loc_cld) {
  const char *path = reference_path();
  h5lite_file *f = open_or_die(path);
  const int nblocks = ngptot / nproma + (ngptot % nproma ? 1 : 0);
  double *scratch = malloc(sizeof(double) * (size_t)nclv * (nlev + 1) * nlon);
  if (!scratch) { fprintf(stderr, "load_reference: out of memory\n"); exit(2); }
#define R(name, rows, field) load_and_expand(f, path, name, rows, nlon, nproma, nblocks, field, scratch)
  R("PLUDE", nlev, plude); R("PCOVPTOT", nlev, pcovptot); R("PRAINFRAC_TOPRFZ", 1, prainfrac_toprfz);
  R("PFSQLF", nlev + 1, pfsqlf); R("PFSQIF", nlev + 1, pfsqif); R("PFCQLNG", nlev + 1, pfcqlng); R("PFCQNNG", nlev + 1, pfcqnng);
  R("PFSQRF", nlev + 1, pfsqrf); R("PFSQSF", nlev + 1, pfsqsf); R("PFCQRNG", nlev + 1, pfcqrng); R("PFCQSNG", nlev + 1, pfcqsng);
  R("PFSQLTUR", nlev + 1, pfsqltur); R("PFSQITUR", nlev + 1, pfsqitur); R("PFPLSL", nlev + 1, pfplsl); R("PFPLSN", nlev + 1, pfplsn);
  R("PFHPSL", nlev + 1, pfhpsl); R("PFHPSN", nlev + 1, pfhpsn);
  R("TENDENCY_LOC_A", nlev, tend_loc_a); R("TENDENCY_LOC_Q", nlev, tend_loc_q); R("TENDENCY_LOC_T", nlev, tend_loc_t);
  R("TENDENCY_LOC_CLD", nclv * nlev, tend_loc_cld);
#undef R
  free(scratch);
  h5lite_close(f);
}
