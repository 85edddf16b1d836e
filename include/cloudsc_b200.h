/*
 * cloudsc_b200.h -- C ABI of the B200-native CLOUDSC hot path (libcloudsc_b200.so).
 *
 * Plain C: pointers, ints and doubles only; no CUDA or torch types cross this boundary
 * (a stream is passed as an opaque `void*` = cudaStream_t, NULL = default stream).
 * Every entry point returns 0 on success and a non-zero code otherwise
 * (CLOUDSC_B200_E* below, or a cudaError_t value); nothing here ever calls exit().
 * `cloudsc_b200_last_error()` returns a human-readable description of the last failure
 * on the calling thread.
 *
 * What each entry point replaces in the reference (paths relative to the dwarf-p-cloudsc
 * checkout, src/cloudsc_c/cloudsc/...):
 *
 *   cloudsc_c()                      cloudsc_c.h:18-29 / cloudsc_c.c:19-2587   binary-compatible
 *                                    drop-in: same 61 arguments, host pointers, one NPROMA block,
 *                                    reads the yrecldp/yomcst/yoethf globals, returns 0.
 *   cloudsc_b200_set_params()        the implicit global inputs of cloudsc_c() (cloudsc_c.c:385-391,
 *                                    every yrecldp-> / rg / r2es ... use) snapshotted into
 *                                    __constant__ memory.
 *   cloudsc_b200_launch()            the OpenMP block loop of cloudsc_driver.c:183-229, incl. the
 *                                    zeroing of pcovptot / tendency_loc_cld at :199-200, as ONE
 *                                    kernel launch over a contiguous range of NPROMA blocks of
 *                                    device-resident fields.
 *   cloudsc_b200_run_host()          the same loop for HOST-resident fields: chunked, pinned,
 *                                    multi-stream H2D || kernel || D2H pipeline (successor of
 *                                    src/cloudsc_cuda/cloudsc/cloudsc_driver.cu:346-447).
 *   cloudsc_b200_create()/upload()/  the cudaMalloc / cudaMemcpy block of
 *     download()/destroy()           src/cloudsc_cuda/cloudsc/cloudsc_driver.cu:276-447.
 *
 * Field tables: all bulk entry points take the 56 field pointers of cloudsc_c() as ONE array
 * `void *fields[CLOUDSC_NFIELDS]`, indexed by `enum cloudsc_field` = position in the
 * cloudsc_c() argument list.  Global layout per field is the driver's:
 * [nblocks][shape-of-one-block], NPROMA fastest (cloudsc_driver.c:193-196).
 */
#ifndef CLOUDSC_B200_H
#define CLOUDSC_B200_H

#include <stddef.h>
#include "cloudsc_params.h"

#ifdef __cplusplus
extern "C" {
#endif

/* kind:  L = [klev][nproma], H = [klev+1][nproma], C = [nclv][klev][nproma], S = [nproma] double,
 *        I = [nproma] int
 * use:   R read, W written, B both (in/out), N never touched by the kernel,
 *        A read only when the matching yrecldp->laer* flag is set: PRE_ICE with laericesed (cloudsc_c.c:1561-1566),
 *          PICRIT_AER and PNICE with laericeauto (:1625-1632); PLCRIT_AER / PCCN belong to laerliqautolsp / laerliqcoll,
 *          which only act in the iwarmrain == 1 branch the reference compiles dead (:367,1654,1679) -> EPARAMS */
#define CLOUDSC_FIELDS(X) \
  X(PT, L, R) X(PQ, L, R) \
  X(TENDENCY_CML_T, L, N) X(TENDENCY_CML_Q, L, N) X(TENDENCY_CML_A, L, N) X(TENDENCY_CML_CLD, C, N) \
  X(TENDENCY_TMP_T, L, R) X(TENDENCY_TMP_Q, L, R) X(TENDENCY_TMP_A, L, R) X(TENDENCY_TMP_CLD, C, R) \
  X(TENDENCY_LOC_T, L, W) X(TENDENCY_LOC_Q, L, W) X(TENDENCY_LOC_A, L, W) X(TENDENCY_LOC_CLD, C, W) \
  X(PVFA, L, N) X(PVFL, L, R) X(PVFI, L, R) X(PDYNA, L, N) X(PDYNL, L, N) X(PDYNI, L, N) \
  X(PHRSW, L, R) X(PHRLW, L, R) X(PVERVEL, L, R) X(PAP, L, R) X(PAPH, H, R) X(PLSM, S, R) \
  X(KTYPE, I, R) X(PLU, L, R) X(PLUDE, L, B) X(PSNDE, L, R) X(PMFU, L, R) X(PMFD, L, R) \
  X(PA, L, R) X(PCLV, C, R) X(PSUPSAT, L, R) \
  X(PLCRIT_AER, L, A) X(PICRIT_AER, L, A) X(PRE_ICE, L, A) X(PCCN, L, A) X(PNICE, L, A) \
  X(PCOVPTOT, L, W) X(PRAINFRAC_TOPRFZ, S, W) \
  X(PFSQLF, H, W) X(PFSQIF, H, W) X(PFCQNNG, H, W) X(PFCQLNG, H, W) X(PFSQRF, H, W) \
  X(PFSQSF, H, W) X(PFCQRNG, H, W) X(PFCQSNG, H, W) X(PFSQLTUR, H, W) X(PFSQITUR, H, W) \
  X(PFPLSL, H, W) X(PFPLSN, H, W) X(PFHPSL, H, W) X(PFHPSN, H, W)

enum cloudsc_field {
#define CLOUDSC_X(name, kind, use) CLOUDSC_F_##name,
  CLOUDSC_FIELDS(CLOUDSC_X)
#undef CLOUDSC_X
  CLOUDSC_NFIELDS /* = 56 */
};

enum {
  CLOUDSC_B200_OK = 0,
  CLOUDSC_B200_EPARAMS = -1,   /* parameter set outside what the kernel is specialised for */
  CLOUDSC_B200_EARG = -2,      /* bad argument (NULL field, klev/nproma/ngptot out of range) */
  CLOUDSC_B200_ENOTINIT = -3,  /* set_params() has not been called on this device */
  CLOUDSC_B200_ENOMEM = -4
};

/* Precision of the device arithmetic / device field storage. */
enum { CLOUDSC_B200_FP64 = 0, CLOUDSC_B200_FP32 = 1 };

/* ---- information ------------------------------------------------------------------------ */
const char *cloudsc_b200_version(void);
const char *cloudsc_b200_last_error(void);
/* number of visible CUDA devices, or a negative CLOUDSC_B200_E* / -cudaError code */
int cloudsc_b200_device_count(void);
/* name/kind/use of field `f` as one-letter codes from the table above (for bindings) */
const char *cloudsc_b200_field_name(int f);
char cloudsc_b200_field_kind(int f);
char cloudsc_b200_field_use(int f);
/* elements one NPROMA block of field f holds (e.g. klev*nproma) */
size_t cloudsc_b200_field_block_elems(int f, int nproma, int klev);

/* ---- devices and pinned host memory ------------------------------------------------------- */
/* select the CUDA device of the calling host thread (one handle / one host thread per GPU;
 * successor of the implicit device 0 of src/cloudsc_cuda/cloudsc/cloudsc_driver.cu) */
int cloudsc_b200_set_device(int device);
/* page-locked host memory for the field arrays (what makes H2D/D2H asynchronous; the reference
 * driver uses plain malloc, cloudsc_driver.c:119-179) -- or pin memory the caller already owns */
int cloudsc_b200_host_alloc(void **out, size_t bytes);
/* flags: WRITE_COMBINED = uncached on the CPU side (upload-only buffers: the device reads them without snooping);
 * HUGEPAGES = 2 MB pages (transparent huge pages) first-touched by the calling thread, then page-locked: 512x fewer
 * IOMMU translations per byte and NUMA placement by the caller's affinity.  Free either kind with host_free(). */
enum { CLOUDSC_B200_HOST_WRITE_COMBINED = 1, CLOUDSC_B200_HOST_HUGEPAGES = 2 };
int cloudsc_b200_host_alloc_ex(void **out, size_t bytes, unsigned flags);
int cloudsc_b200_host_free(void *p);
int cloudsc_b200_host_register(void *p, size_t bytes);
int cloudsc_b200_host_unregister(void *p);

/* ---- parameters -> __constant__ memory of the current device ---------------------------- */
int cloudsc_b200_set_params(const struct TECLDP *ydecldp, const struct cloudsc_consts *consts);
/* convenience for C callers that use the reference's globals (yrecldp, rg, r2es, ...) */
int cloudsc_b200_set_params_from_globals(void);

/* ---- device-resident launch ------------------------------------------------------------- */
/*
 * Run CLOUDSC on columns [0, ngptot) held in device memory as `nblocks = ceil(ngptot/nproma)`
 * blocks.  `dev_fields[f]` are device pointers (may be NULL for use N / A fields).
 * precision = CLOUDSC_B200_FP64: fields are double; FP32: fields are float (KTYPE stays int).
 * Asynchronous on `stream`.  PLUDE is in/out (cloudsc_c.c:970,980).
 */
int cloudsc_b200_launch(void *const dev_fields[CLOUDSC_NFIELDS], int ngptot, int nproma, int klev,
                        double ptsphy, int precision, void *stream);

/* Kernel variant selection for sweeps/profiling: 0 = default (auto). See DESIGN.md. */
int cloudsc_b200_set_variant(int variant);
int cloudsc_b200_get_variant(void);
/* registers/thread, spill bytes, static smem, max threads/block, achieved blocks/SM of the
 * kernel the next launch(nproma, precision) would use; out[8] ints */
int cloudsc_b200_kernel_info(int nproma, int precision, int out[8]);
/* number of kernels this process has launched since the library was loaded, per kernel family:
 * out[0] tensor-map TMA kernel, [1] bulk-copy TMA kernel, [2] plain-load kernel, [3] cp.async-ring kernel,
 * [4] plain-load kernel with the aerosol inputs (laericesed / laericeauto) */
int cloudsc_b200_launch_counts(long long out[5]);
/* Self-test of the arithmetic primitives of the product kernels on the current device: for each of the n inputs, out[12 i + ..] =
 * { max(x,0), min(x,1), min(x,0.5), max(x,160), max(x,-700) as the kernels evaluate them (high-word integer tests),
 *   exp_fast(clamp(x, -700, 700)), log_fast(|x|), rcp_fast(x), sqrt_fast(|x|), div_fast(1.2345678901234567, x) (x != 0),
 *   pow_fast(|x|, 0.666), pow_fast(|x|, 2.47) }.  Device pointers.  tests/ check the first five
 * bit for bit against the plain compares on zeros of both signs, denormals and the neighbours of every threshold. */
int cloudsc_b200_selftest_primitives(const double *x, int n, double *out, void *stream);

/* ---- host-resident, end-to-end ---------------------------------------------------------- */
/*
 * Same computation with HOST pointers (pageable or pinned), all blocks: stages chunks of
 * `chunk_blocks` NPROMA-blocks through pinned buffers on `nstreams` streams (0 = defaults),
 * uploading only fields with use R/B and downloading only W/B.  Synchronous.
 * If `stats` != NULL it receives {h2d_bytes, d2h_bytes, kernel_ms_sum, wall_ms}.
 */
int cloudsc_b200_run_host(void *const host_fields[CLOUDSC_NFIELDS], int ngptot, int nproma, int klev,
                          double ptsphy, int chunk_blocks, int nstreams, double stats[4]);
/* the same with the precision of the host arrays AND of the device arithmetic: CLOUDSC_B200_FP64 (= run_host) or
 * CLOUDSC_B200_FP32 (float host fields, KTYPE int).
 * What travels: of every R/B field only the rows the kernel reads (species 0..3 of PCLV / TENDENCY_TMP_CLD; no rows
 * above level ncldtop-1 of PQ, TENDENCY_TMP_Q, PSUPSAT, PSNDE, PHRSW, PHRLW, PVERVEL, PMFU, PMFD, PAP, PLU), of every
 * W/B field the rows it writes (species 0..3 of TENDENCY_LOC_CLD; species 4 is zeroed on the host, cloudsc_driver.c:200).
 * Of a ragged last block only the first ngptot - (nblocks-1)*nproma columns are written back; the padding columns of
 * the caller's output arrays stay untouched, except those of PCOVPTOT / TENDENCY_LOC_CLD, which get the zeros the
 * reference driver writes (cloudsc_driver.c:199-200). */
int cloudsc_b200_run_host_ex(void *const host_fields[CLOUDSC_NFIELDS], int ngptot, int nproma, int klev,
                             double ptsphy, int precision, int chunk_blocks, int nstreams, double stats[4]);
/* run_host keeps its streams and device staging arenas between calls (allocation costs more than
 * the copies); this frees them on every device */
int cloudsc_b200_release_cache(void);

/* ---- handle API: resident device state for a fixed geometry (one handle per GPU) --------- */
/*
 * The state owns ONE device arena in which the fields of a kind lie at a constant spacing.  That lets
 * step() describe the inputs with four tensor maps and fetch a level's 27 input rows with 4 TMA requests
 * instead of 27; cloudsc_b200_launch() on caller-owned pointers cannot assume that.
 */
typedef struct cloudsc_b200_state cloudsc_b200_state;
int cloudsc_b200_create(int ngptot, int nproma, int klev, int precision, cloudsc_b200_state **out);
int cloudsc_b200_destroy(cloudsc_b200_state *st);
/* H2D of every R/B field (host arrays of the state's element type: double, or float for FP32 states; KTYPE int) */
int cloudsc_b200_upload(cloudsc_b200_state *st, void *const host_fields[CLOUDSC_NFIELDS], void *stream);
/* D2H of every W/B field */
int cloudsc_b200_download(cloudsc_b200_state *st, void *const host_fields[CLOUDSC_NFIELDS], void *stream);
/* launch over blocks [first_block, first_block+nblocks) of the state; nblocks<=0 -> all */
int cloudsc_b200_step(cloudsc_b200_state *st, double ptsphy, int first_block, int nblocks, void *stream);
/* how step() schedules the 32-column tiles of a launch: mode 0 = one CTA per tile; mode 1 = persistent warps that run
 * the upper (bandwidth-bound) and lower (latency-bound) parts of their tiles interleaved, cut at level `ksplit`, with
 * about `target_b` of an SM's 12 warps in a lower part at any time (<= 0: defaults); mode 2 = persistent warps taking whole
 * tiles from a queue; mode 3 (the default of an FP64 state) = one launch whose first `target_b` x #SM blocks sweep their
 * tile only down to level `ksplit` and whose last blocks finish those tiles, so that the launch ends on half tiles.
 * Bit-identical results on the shipped data; rounding-level differences between kernels elsewhere. */
int cloudsc_b200_set_schedule(cloudsc_b200_state *st, int mode, int ksplit, int target_b);
/* out = { mode, ksplit, target_b as set (0 = defaults), the schedule a step() over the WHOLE state actually gets }: a launch
 * too small for the schedule (modes 1, 2: fewer than two rounds of tiles per resident warp; mode 3: every tile resident at
 * once, i.e. <= 56 832 columns on 148 SMs in FP64), an aerosol run or a state without tensor maps keeps one CTA per tile
 * (0) whatever the mode.  A state is not re-entrant under modes 1-3: one step() of a state in flight at a time (steps on one
 * stream are; the schedules keep per-state scratch), and a step() captured into a CUDA graph should use mode 0. */
int cloudsc_b200_get_schedule(cloudsc_b200_state *st, int out[4]);
/* restore PLUDE from the pristine copy kept at upload() (the kernel scales it in place) */
int cloudsc_b200_reset_inout(cloudsc_b200_state *st, void *stream);
/* kernel_info of the kernel step() launches for this state (the arena-backed tensor-map TMA kernel when
 * NPROMA is a multiple of 32) */
int cloudsc_b200_state_kernel_info(cloudsc_b200_state *st, int out[8]);
/* device pointer of field f inside the state (for zero-copy consumers) */
void *cloudsc_b200_field_ptr(cloudsc_b200_state *st, int f);
int cloudsc_b200_sync(void *stream);

/* ---- binary-compatible drop-in for the reference kernel (host pointers, one block) -------- */
/* Re-entrant: called concurrently from the OpenMP threads of cloudsc_driver.c:191-217, each thread stages its block
 * through thread-local device buffers on its own stream.  The buffers are freed when the thread exits, or by
 * cloudsc_b200_release_thread() from that thread. */
int cloudsc_b200_release_thread(void);
int cloudsc_c(int kidia, int kfdia, int klon, int klev, double ptsphy,
  double *v_pt, double *v_pq,
  double *v_tendency_cml_t, double *v_tendency_cml_q, double *v_tendency_cml_a, double *v_tendency_cml_cld,
  double *v_tendency_tmp_t, double *v_tendency_tmp_q, double *v_tendency_tmp_a, double *v_tendency_tmp_cld,
  double *v_tendency_loc_t, double *v_tendency_loc_q, double *v_tendency_loc_a, double *v_tendency_loc_cld,
  double *v_pvfa, double *v_pvfl, double *v_pvfi, double *v_pdyna, double *v_pdynl, double *v_pdyni,
  double *v_phrsw, double *v_phrlw, double *v_pvervel, double *v_pap, double *v_paph, double *v_plsm,
  int *v_ktype, double *v_plu, double *v_plude, double *v_psnde, double *v_pmfu, double *v_pmfd,
  double *v_pa, double *v_pclv, double *v_psupsat, double *v_plcrit_aer, double *v_picrit_aer,
  double *v_pre_ice, double *v_pccn, double *v_pnice, double *v_pcovptot, double *v_prainfrac_toprfz,
  double *v_pfsqlf, double *v_pfsqif, double *v_pfcqnng, double *v_pfcqlng, double *v_pfsqrf,
  double *v_pfsqsf, double *v_pfcqrng, double *v_pfcqsng, double *v_pfsqltur, double *v_pfsqitur,
  double *v_pfplsl, double *v_pfplsn, double *v_pfhpsl, double *v_pfhpsn);

#ifdef __cplusplus
}
#endif
#endif /* CLOUDSC_B200_H */
