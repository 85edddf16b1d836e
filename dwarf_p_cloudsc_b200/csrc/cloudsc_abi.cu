/*
 * cloudsc_abi.cu -- the extern "C" boundary of libcloudsc_b200.so (declared in include/cloudsc_b200.h).
 *
 * Host-side plumbing only: parameter snapshot into __constant__ memory, device-resident launch,
 * the pinned/chunked/multi-stream host pipeline, the resident-state handle API and the
 * binary-compatible cloudsc_c() shim.  There is NO host arithmetic path in this library: every
 * entry point either runs the sm_100a kernels of cloudsc_kernels.cu or returns an error.
 */
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include <sys/mman.h>
#include <time.h>

#include <cuda_runtime.h>

#include "cloudsc_b200.h"
#include "cloudsc_fill.h"
#include "cloudsc_kernels.cuh"

using namespace cloudsc;

namespace {

thread_local char t_err[512] = "";

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
  return code;
}
int cuda_fail(cudaError_t e, const char *what) {
  return fail((int)e, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}
#define CU(call)                                         \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call);  \
  } while (0)

const char k_kind[CLOUDSC_NFIELDS + 1] = {
#define CLOUDSC_X(name, kind, use) #kind[0],
    CLOUDSC_FIELDS(CLOUDSC_X)
#undef CLOUDSC_X
    0};
const char k_use[CLOUDSC_NFIELDS + 1] = {
#define CLOUDSC_X(name, kind, use) #use[0],
    CLOUDSC_FIELDS(CLOUDSC_X)
#undef CLOUDSC_X
    0};
const char *const k_name[CLOUDSC_NFIELDS] = {
#define CLOUDSC_X(name, kind, use) #name,
    CLOUDSC_FIELDS(CLOUDSC_X)
#undef CLOUDSC_X
};

/* `aer`: bit 0 = yrecldp->laericesed (PRE_ICE is read), bit 1 = laericeauto (PICRIT_AER, PNICE are read); the other
 * use-A fields belong to switches the kernel refuses */
inline bool is_aer_in(int f, int aer) {
  return (f == CLOUDSC_F_PRE_ICE && (aer & 1)) || ((f == CLOUDSC_F_PICRIT_AER || f == CLOUDSC_F_PNICE) && (aer & 2));
}
inline bool is_in(int f, int aer = 0) { return k_use[f] == 'R' || k_use[f] == 'B' || is_aer_in(f, aer); }
inline bool is_out(int f) { return k_use[f] == 'W' || k_use[f] == 'B'; }
inline bool is_live(int f, int aer = 0) { return is_in(f, aer) || is_out(f); }
inline size_t elem_size(int f, int precision) {
  return k_kind[f] == 'I' ? sizeof(int) : (precision == CLOUDSC_B200_FP32 ? sizeof(float) : sizeof(double));
}
inline size_t rows_of(int f, int klev) {
  switch (k_kind[f]) {
    case 'L': return (size_t)klev;
    case 'H': return (size_t)klev + 1;
    case 'C': return (size_t)5 * klev;
    default: return 1;
  }
}

/* which devices have had set_params() */
std::mutex g_mu;
bool g_params_set[64] = {false};
std::atomic<int> g_variant{0};

int current_device_ready() {
  int dev = 0;
  CU(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_mu);
  if (dev < 0 || dev >= 64 || !g_params_set[dev])
    return fail(CLOUDSC_B200_ENOTINIT, "cloudsc_b200_set_params() has not been called on device %d", dev);
  return 0;
}

int check_geom(int ngptot, int nproma, int klev) {
  if (ngptot <= 0 || nproma <= 0 || klev < 2) return fail(CLOUDSC_B200_EARG, "bad geometry ngptot=%d nproma=%d klev=%d", ngptot, nproma, klev);
  return 0;
}

int check_fields(void *const f[CLOUDSC_NFIELDS], int aer) {
  if (!f) return fail(CLOUDSC_B200_EARG, "field table is NULL");
  for (int i = 0; i < CLOUDSC_NFIELDS; ++i)
    if (is_live(i, aer) && !f[i]) return fail(CLOUDSC_B200_EARG, "field %s is NULL", k_name[i]);
  return 0;
}

inline size_t up256(size_t n) { return (n + 255) / 256 * 256; }

/*
 * Layout of a device arena holding `nblocks` NPROMA blocks of every live field: the three tensor-map groups first,
 * each with a constant spacing between its fields (so that ONE tensor-map request fetches the rows of all of them,
 * cloudsc_kernels.cuh), then paph, then the rest.  Used by the resident state and by the staging lanes of run_host.
 */
struct ArenaLayout {
  size_t off[CLOUDSC_NFIELDS], bytes[CLOUDSC_NFIELDS], total;
};
ArenaLayout arena_layout(int nblocks, int nproma, int klev, int precision, int aer) {
  ArenaLayout a;
  bool placed[CLOUDSC_NFIELDS] = {false};
  a.total = 0;
  for (int f = 0; f < CLOUDSC_NFIELDS; ++f) {
    a.off[f] = 0;
    a.bytes[f] = is_live(f, aer) ? (size_t)nblocks * rows_of(f, klev) * nproma * elem_size(f, precision) : 0;
  }
  auto place = [&](int f) { a.off[f] = a.total; a.total += up256(a.bytes[f]); placed[f] = true; };
  for (int f : kArenaGroupA) place(f);
  for (int f : kArenaGroupB) place(f);
  for (int f : kArenaGroupC) place(f);
  place(CLOUDSC_F_PAPH);
  for (int f = 0; f < CLOUDSC_NFIELDS; ++f)
    if (is_live(f, aer) && !placed[f]) place(f);
  if (a.total == 0) a.total = 256;
  return a;
}

}  // namespace

/* ============================================================================================== */
extern "C" {

const char *cloudsc_b200_version(void) { return "cloudsc_b200 0.2 (sm_100a)"; }
const char *cloudsc_b200_last_error(void) { return t_err; }

int cloudsc_b200_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cuda_fail(e, "cudaGetDeviceCount"); return -(int)e; }
  return n;
}

const char *cloudsc_b200_field_name(int f) { return (f >= 0 && f < CLOUDSC_NFIELDS) ? k_name[f] : ""; }
char cloudsc_b200_field_kind(int f) { return (f >= 0 && f < CLOUDSC_NFIELDS) ? k_kind[f] : 0; }
char cloudsc_b200_field_use(int f) { return (f >= 0 && f < CLOUDSC_NFIELDS) ? k_use[f] : 0; }
size_t cloudsc_b200_field_block_elems(int f, int nproma, int klev) {
  if (f < 0 || f >= CLOUDSC_NFIELDS) return 0;
  return rows_of(f, klev) * (size_t)nproma;
}

int cloudsc_b200_set_device(int device) {
  CU(cudaSetDevice(device));
  return 0;
}

int cloudsc_b200_host_alloc(void **out, size_t bytes) {
  if (!out) return fail(CLOUDSC_B200_EARG, "host_alloc: out is NULL");
  *out = nullptr;
  CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
  return 0;
}

namespace {
/* blocks handed out by host_alloc_ex(CLOUDSC_B200_HOST_HUGEPAGES): malloc'ed + registered, not cudaHostAlloc'ed */
std::mutex g_reg_mu;
std::vector<void *> g_registered;
}  // namespace

int cloudsc_b200_host_alloc_ex(void **out, size_t bytes, unsigned flags) {
  if (!out) return fail(CLOUDSC_B200_EARG, "host_alloc_ex: out is NULL");
  *out = nullptr;
  if (!bytes) bytes = 1;
  if (flags & CLOUDSC_B200_HOST_HUGEPAGES) {
    /* 2 MB-aligned anonymous memory, transparent huge pages requested, touched by THIS thread (first touch places
       the pages on the caller's NUMA node), then page-locked */
    const size_t align = (size_t)2 << 20, n = (bytes + align - 1) / align * align;
    void *p = aligned_alloc(align, n);
    if (!p) return fail(CLOUDSC_B200_ENOMEM, "host_alloc_ex: out of host memory (%zu bytes)", n);
    madvise(p, n, MADV_HUGEPAGE);
    for (size_t o = 0; o < n; o += 4096) ((volatile char *)p)[o] = 0;
    cudaError_t e = cudaHostRegister(p, n, cudaHostRegisterPortable);
    if (e != cudaSuccess) { free(p); return cuda_fail(e, "cudaHostRegister"); }
    std::lock_guard<std::mutex> lk(g_reg_mu);
    g_registered.push_back(p);
    *out = p;
    return 0;
  }
  unsigned f = cudaHostAllocPortable;
  if (flags & CLOUDSC_B200_HOST_WRITE_COMBINED) f |= cudaHostAllocWriteCombined;
  CU(cudaHostAlloc(out, bytes, f));
  return 0;
}

int cloudsc_b200_host_free(void *p) {
  if (!p) return 0;
  {
    std::lock_guard<std::mutex> lk(g_reg_mu);
    for (size_t i = 0; i < g_registered.size(); ++i)
      if (g_registered[i] == p) {
        g_registered.erase(g_registered.begin() + (long)i);
        cudaHostUnregister(p);
        free(p);
        return 0;
      }
  }
  CU(cudaFreeHost(p));
  return 0;
}

int cloudsc_b200_host_register(void *p, size_t bytes) {
  if (!p || !bytes) return 0;
  CU(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  return 0;
}

int cloudsc_b200_host_unregister(void *p) {
  if (p) CU(cudaHostUnregister(p));
  return 0;
}

namespace {
/* what set_params() last uploaded to each device: identical values are not uploaded again, so that the per-thread
 * first calls of the cloudsc_c() shim do not rewrite __constant__ memory under kernels other threads have in flight */
CloudscConst<double> g_uploaded[64];
}  // namespace

int cloudsc_b200_set_params(const struct TECLDP *p, const struct cloudsc_consts *k) {
  if (!p || !k) return fail(CLOUDSC_B200_EARG, "set_params: NULL argument");
  if (const char *why = check_params(p)) return fail(CLOUDSC_B200_EPARAMS, "unsupported parameter set: %s", why);
  CloudscConst<double> c64;
  CloudscConst<float> c32;
  memset(&c64, 0, sizeof(c64));
  memset(&c32, 0, sizeof(c32));
  fill_consts<double>(p, k, &c64);
  fill_consts<float>(p, k, &c32);
  int dev = 0;
  CU(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(CLOUDSC_B200_EARG, "device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_params_set[dev] && memcmp(&g_uploaded[dev], &c64, sizeof(c64)) == 0) return 0;
  CU(upload_consts(c64, c32));
  g_uploaded[dev] = c64;
  g_params_set[dev] = true;
  return 0;
}

int cloudsc_b200_set_params_from_globals(void) {
  if (!yrecldp) return fail(CLOUDSC_B200_EARG, "yrecldp is NULL");
  struct cloudsc_consts k;
#define CLOUDSC_X(n) k.n = n;
  CLOUDSC_KERNEL_CONSTS(CLOUDSC_X)
#undef CLOUDSC_X
  return cloudsc_b200_set_params(yrecldp, &k);
}

int cloudsc_b200_set_variant(int variant) {
  if (!variant_valid(variant)) return fail(CLOUDSC_B200_EARG, "unknown kernel variant %d", variant);
  g_variant.store(variant);
  return 0;
}
int cloudsc_b200_get_variant(void) { return g_variant.load(); }

int cloudsc_b200_kernel_info(int nproma, int precision, int out[8]) {
  if (!out) return fail(CLOUDSC_B200_EARG, "kernel_info: out is NULL");
  CU(sweep_kernel_info(nproma, precision, g_variant.load(), out));
  return 0;
}

/* ---------------------------------------------------------------------------------------------- */
/* tensor maps over a regularly spaced field set                                                   */
/* ---------------------------------------------------------------------------------------------- */
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn lookup_encode_tiled() {
  void *p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
    return (EncodeTiledFn)p;
  return nullptr;
}
EncodeTiledFn encode_tiled_fn() {
  static const EncodeTiledFn fn = lookup_encode_tiled();   /* C++11: initialised once, thread-safe */
  return fn;
}
/* rank-R tensor map over `base`: dims/box innermost first, strides (bytes) of dims 1..R-1 */
bool encode_map(CUtensorMap *m, int precision, void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides,
                const cuuint32_t *box) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  return fn(m, precision == CLOUDSC_B200_FP32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

/*
 * The tensor-map kernel (4 TMA requests per level instead of 27) needs the input fields of each group at a constant
 * spacing: dev[group[i]] == dev[group[0]] + i * spacing.  That holds for the library's own arenas (resident state,
 * run_host lanes) and for any caller that carves its fields out of one allocation in the group order
 * (INTEGRATION.md); caller-owned fields from separate cudaMalloc calls normally do not qualify and run the
 * bulk-copy kernel.  `nblocks` = blocks the fields are allocated for.
 */
bool regular_spacing(void *const dev[CLOUDSC_NFIELDS], const int *group, int n, size_t min_bytes, size_t *spacing) {
  const char *p0 = (const char *)dev[group[0]];
  if (!p0 || ((uintptr_t)p0 & 15)) return false;
  if (n == 1) { *spacing = up256(min_bytes); return true; }
  const char *p1 = (const char *)dev[group[1]];
  if (!p1 || p1 <= p0) return false;
  const size_t sp = (size_t)(p1 - p0);
  if (sp < min_bytes || (sp & 15)) return false;
  for (int i = 2; i < n; ++i)
    if ((const char *)dev[group[i]] != p0 + (size_t)i * sp) return false;
  *spacing = sp;
  return true;
}

bool build_tensor_maps(void *const dev[CLOUDSC_NFIELDS], int nblocks, int nproma, int klev, int precision, TensorMaps *tm) {
  if (nproma % 32 != 0 || nblocks <= 0) return false;
  const size_t es = elem_size(CLOUDSC_F_PT, precision);
  const size_t bytesL = (size_t)nblocks * klev * nproma * es, bytesC = 5 * bytesL;
  size_t SA = 0, SB = 0, SC = 0;
  if (!regular_spacing(dev, kArenaGroupA, 15, bytesL, &SA)) return false;
  if (!regular_spacing(dev, kArenaGroupB, 3, bytesL, &SB)) return false;
  if (!regular_spacing(dev, kArenaGroupC, 2, bytesC, &SC)) return false;
  if (!dev[CLOUDSC_F_PAPH] || ((uintptr_t)dev[CLOUDSC_F_PAPH] & 15)) return false;
  const cuuint64_t np = (cuuint64_t)nproma, kl = (cuuint64_t)klev, nb = (cuuint64_t)nblocks;
  const cuuint64_t rowb = np * es;
  const cuuint64_t dA[4] = {np, kl, nb, 15}, sA[3] = {rowb, kl * rowb, SA};
  const cuuint32_t bA[4] = {32, 1, 1, 15}, bAtop[4] = {32, 1, 1, 7};
  const cuuint64_t dB[4] = {np, kl, nb, 3}, sB[3] = {rowb, kl * rowb, SB};
  const cuuint32_t bB[4] = {32, 1, 1, 3};
  const cuuint64_t dC[5] = {np, kl, 5, nb, 2}, sC[4] = {rowb, kl * rowb, 5 * kl * rowb, SC};
  const cuuint32_t bC[5] = {32, 1, 4, 1, 2};
  const cuuint64_t dH[3] = {np, kl + 1, nb}, sH[2] = {rowb, (kl + 1) * rowb};
  const cuuint32_t bH[3] = {32, 1, 1};
  return encode_map(&tm->a, precision, dev[kArenaGroupA[0]], 4, dA, sA, bA) &&
         encode_map(&tm->a_top, precision, dev[kArenaGroupA[0]], 4, dA, sA, bAtop) &&
         encode_map(&tm->b, precision, dev[kArenaGroupB[0]], 4, dB, sB, bB) &&
         encode_map(&tm->c, precision, dev[kArenaGroupC[0]], 5, dC, sC, bC) &&
         encode_map(&tm->h, precision, dev[CLOUDSC_F_PAPH], 3, dH, sH, bH);
}

/* small per-process cache for cloudsc_b200_launch(): encoding four maps costs ~10 us, a launch ~5 */
struct TmapCacheEntry {
  void *key[4];
  int device, nblocks, nproma, klev, precision;
  bool ok;
  TensorMaps tm;
};
std::mutex g_tmap_mu;
TmapCacheEntry g_tmap_cache[8];
int g_tmap_next = 0, g_tmap_used = 0;

bool cached_tensor_maps(void *const dev[CLOUDSC_NFIELDS], int device, int nblocks, int nproma, int klev, int precision, TensorMaps *out) {
  void *key[4] = {dev[kArenaGroupA[0]], dev[kArenaGroupA[14]], dev[kArenaGroupC[1]], dev[CLOUDSC_F_PAPH]};
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  for (int i = 0; i < g_tmap_used; ++i) {
    TmapCacheEntry &e = g_tmap_cache[i];
    if (memcmp(e.key, key, sizeof(key)) == 0 && e.device == device && e.nblocks == nblocks && e.nproma == nproma && e.klev == klev &&
        e.precision == precision) {
      /* the cheap part of the key can collide after a free/realloc with another layout: re-check the spacing */
      size_t sp;
      if (e.ok && !(regular_spacing(dev, kArenaGroupA, 15, 0, &sp) && regular_spacing(dev, kArenaGroupB, 3, 0, &sp) &&
                    regular_spacing(dev, kArenaGroupC, 2, 0, &sp)))
        break;
      if (e.ok) *out = e.tm;
      return e.ok;
    }
  }
  TmapCacheEntry &e = g_tmap_cache[g_tmap_next];
  g_tmap_next = (g_tmap_next + 1) % 8;
  if (g_tmap_used < 8) ++g_tmap_used;
  memcpy(e.key, key, sizeof(key));
  e.device = device; e.nblocks = nblocks; e.nproma = nproma; e.klev = klev; e.precision = precision;
  e.ok = build_tensor_maps(dev, nblocks, nproma, klev, precision, &e.tm);
  if (e.ok) *out = e.tm;
  return e.ok;
}
}  // namespace

int cloudsc_b200_launch(void *const dev_fields[CLOUDSC_NFIELDS], int ngptot, int nproma, int klev, double ptsphy,
                        int precision, void *stream) {
  if (int rc = check_geom(ngptot, nproma, klev)) return rc;
  if (precision != CLOUDSC_B200_FP64 && precision != CLOUDSC_B200_FP32) return fail(CLOUDSC_B200_EARG, "bad precision %d", precision);
  if (int rc = current_device_ready()) return rc;
  const int aer = sweep_aerosol_mask();
  if (int rc = check_fields(dev_fields, aer)) return rc;
  FieldTable ft;
  for (int i = 0; i < CLOUDSC_NFIELDS; ++i) ft.p[i] = dev_fields[i];
  LaunchGeom g{0, ngptot, nproma, klev};
  const int variant = g_variant.load();
  if (variant == VARIANT_AUTO && !aer && nproma % 32 == 0) {
    int dev = 0;
    CU(cudaGetDevice(&dev));
    TensorMaps tm;
    if (cached_tensor_maps(dev_fields, dev, (ngptot + nproma - 1) / nproma, nproma, klev, precision, &tm)) {
      CU(launch_sweep_tmap(ft, tm, g, ptsphy, precision, (cudaStream_t)stream));
      return 0;
    }
  }
  CU(launch_sweep(ft, g, ptsphy, precision, variant, (cudaStream_t)stream));
  return 0;
}

/* kernels launched by this process since the library was loaded, per family:
   {tensor-map TMA, bulk-copy TMA, plain loads, cp.async ring, plain loads with aerosol inputs} */
int cloudsc_b200_selftest_primitives(const double *x, int n, double *out, void *stream) {
  if (n < 0 || (n > 0 && (!x || !out))) return fail(CLOUDSC_B200_EARG, "selftest_primitives: bad arguments");
  CU(launch_selftest_primitives(x, n, out, (cudaStream_t)stream));
  return 0;
}

int cloudsc_b200_launch_counts(long long out[5]) {
  if (!out) return fail(CLOUDSC_B200_EARG, "launch_counts: out is NULL");
  static_assert(KERNEL_FAMILIES == 5, "update cloudsc_b200_launch_counts");
  launch_counts(out);
  return 0;
}

int cloudsc_b200_sync(void *stream) {
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

/* ---------------------------------------------------------------------------------------------- */
/* resident state                                                                                 */
/* ---------------------------------------------------------------------------------------------- */
struct cloudsc_b200_state {
  int ngptot, nproma, klev, nblocks, precision, device;
  int aer;                      /* aerosol mask the state was created for (which use-A fields it holds) */
  void *dev[CLOUDSC_NFIELDS];   /* pointers into `arena` */
  void *plude0;                 /* pristine copy of the in/out field */
  size_t bytes[CLOUDSC_NFIELDS];
  char *arena;                  /* ONE allocation: fields of a kind at a constant spacing (see TensorMaps) */
  TensorMaps tm;
  bool has_tm;                  /* tensor maps encoded: step() uses the 4-requests-per-level TMA kernel */
  /* persistent schedule (cloudsc_b200_set_schedule): scratch for the state at the cut + per-SM counters */
  int sched_mode;               /* 0: one CTA per 32-column tile, 1: persistent warps, upper/lower parts interleaved */
  PersistSchedule sched;
};

namespace {
int state_on_current_device(const cloudsc_b200_state *st, const char *what);
}

int cloudsc_b200_create(int ngptot, int nproma, int klev, int precision, cloudsc_b200_state **out) {
  if (!out) return fail(CLOUDSC_B200_EARG, "create: out is NULL");
  *out = nullptr;
  if (int rc = check_geom(ngptot, nproma, klev)) return rc;
  if (precision != CLOUDSC_B200_FP64 && precision != CLOUDSC_B200_FP32) return fail(CLOUDSC_B200_EARG, "bad precision %d", precision);
  cloudsc_b200_state *st = (cloudsc_b200_state *)calloc(1, sizeof(*st));
  if (!st) return fail(CLOUDSC_B200_ENOMEM, "out of host memory");
  st->ngptot = ngptot; st->nproma = nproma; st->klev = klev; st->precision = precision;
  st->nblocks = (ngptot + nproma - 1) / nproma;
  st->aer = sweep_aerosol_mask();   /* 0 before set_params(); step() refuses aerosol switches the state has no fields for */
  cudaGetDevice(&st->device);

  const ArenaLayout lay = arena_layout(st->nblocks, nproma, klev, precision, st->aer);
  for (int f = 0; f < CLOUDSC_NFIELDS; ++f) st->bytes[f] = lay.bytes[f];
  cudaError_t e = cudaMalloc(&st->arena, lay.total);
  if (e != cudaSuccess) { cloudsc_b200_destroy(st); return cuda_fail(e, "cudaMalloc(arena)"); }
  for (int f = 0; f < CLOUDSC_NFIELDS; ++f) st->dev[f] = is_live(f, st->aer) ? st->arena + lay.off[f] : nullptr;
  e = cudaMalloc(&st->plude0, st->bytes[CLOUDSC_F_PLUDE]);
  if (e != cudaSuccess) { cloudsc_b200_destroy(st); return cuda_fail(e, "cudaMalloc(plude0)"); }
  /* tensor maps (whole 32-column rows: NPROMA must be a multiple of 32) */
  st->has_tm = build_tensor_maps(st->dev, st->nblocks, nproma, klev, precision, &st->tm);
  /* Default schedule of the FP64 state: the split-tail launch (mode 3, cloudsc_kernels.cu) -- 5 % faster than one CTA per tile
     at 163 840 columns (profiles/r2_split_tail.txt), bit-identical on the shipped data; launches too small for it, aerosol runs
     and the FP32 state (not measured) keep one CTA per tile.  Its scratch is small; if it cannot be had, mode 0 stays. */
  if (st->has_tm && precision == CLOUDSC_B200_FP64 && !st->aer) {
    const int ntiles = (st->nblocks * st->nproma + 31) / 32;
    if (cudaMalloc(&st->sched.carry, persist_carry_bytes(ntiles, st->precision)) == cudaSuccess &&
        cudaMalloc((void **)&st->sched.sm_b, kSchedInts * sizeof(int)) == cudaSuccess &&
        cudaMemset(st->sched.sm_b, 0, kSchedInts * sizeof(int)) == cudaSuccess &&
        cudaStreamSynchronize(cudaStreamLegacy) == cudaSuccess) {   /* (a step() on a non-blocking stream must find the flags zeroed) */
      st->sched.carry_tiles = ntiles;
      st->sched_mode = st->sched.mode = 3;
      st->sched.ksplit = st->sched.target_b = 0;
    } else {
      (void)cudaGetLastError();
      if (st->sched.carry) { cudaFree(st->sched.carry); st->sched.carry = nullptr; }
      if (st->sched.sm_b) { cudaFree(st->sched.sm_b); st->sched.sm_b = nullptr; }
    }
  }
  *out = st;
  return 0;
}

int cloudsc_b200_destroy(cloudsc_b200_state *st) {
  if (!st) return 0;
  if (st->arena) cudaFree(st->arena);
  if (st->plude0) cudaFree(st->plude0);
  if (st->sched.carry) cudaFree(st->sched.carry);
  if (st->sched.sm_b) cudaFree(st->sched.sm_b);
  free(st);
  return 0;
}

/*
 * How step() walks the tiles of a launch ("persistent kernel"):
 *   mode 0  one one-warp CTA per 32-column tile, every tile swept top to bottom in one go
 *   mode 2  streaming: 12 persistent warps per SM taking tiles from a queue, the input ring kept full across tiles
 *   mode 1  persistent: 12 warps per SM, each running its tiles in two pieces (levels [0, ksplit) and [ksplit, klev))
 *           and choosing the next piece so that about target_b of the SM's warps are in a lower piece at any time
 *   mode 3  split tail (the default of an FP64 state): one launch whose first `target_b` x #SM blocks sweep their tile only down
 *           to level `ksplit` and park the state, and whose LAST blocks finish those tiles -- the launch ends on half tiles
 * ksplit <= 0 -> 3*klev/5; target_b <= 0 -> 7 of 12 (mode 1), 5 of 12 tiles per SM (mode 3).  Bit-identical in every mode.
 */
int cloudsc_b200_set_schedule(cloudsc_b200_state *st, int mode, int ksplit, int target_b) {
  if (!st) return fail(CLOUDSC_B200_EARG, "set_schedule: NULL state");
  if (mode < 0 || mode > 3) return fail(CLOUDSC_B200_EARG, "set_schedule: unknown mode %d", mode);
  if (ksplit >= st->klev) return fail(CLOUDSC_B200_EARG, "set_schedule: ksplit must be below klev");
  if (int rc = state_on_current_device(st, "set_schedule")) return rc;
  if (mode >= 1 && st->has_tm && !st->sched.carry) {
    const int ntiles = (st->nblocks * st->nproma + 31) / 32;
    CU(cudaMalloc(&st->sched.carry, persist_carry_bytes(ntiles, st->precision)));
    CU(cudaMalloc((void **)&st->sched.sm_b, kSchedInts * sizeof(int)));
    CU(cudaMemset(st->sched.sm_b, 0, kSchedInts * sizeof(int)));
    CU(cudaStreamSynchronize(cudaStreamLegacy));
    st->sched.carry_tiles = ntiles;
  }
  st->sched_mode = mode;
  st->sched.mode = mode;
  st->sched.ksplit = ksplit;
  st->sched.target_b = target_b;
  return 0;
}

namespace {
int state_on_current_device(const cloudsc_b200_state *st, const char *what) {
  int dev = -1;
  CU(cudaGetDevice(&dev));
  if (dev != st->device)
    return fail(CLOUDSC_B200_EARG, "%s: the state lives on device %d but the calling thread's current device is %d", what, st->device, dev);
  return 0;
}
}  // namespace

int cloudsc_b200_upload(cloudsc_b200_state *st, void *const host[CLOUDSC_NFIELDS], void *stream) {
  if (!st || !host) return fail(CLOUDSC_B200_EARG, "upload: NULL argument");
  if (int rc = state_on_current_device(st, "upload")) return rc;
  for (int f = 0; f < CLOUDSC_NFIELDS; ++f) {
    if (!is_in(f, st->aer)) continue;
    if (!host[f]) return fail(CLOUDSC_B200_EARG, "upload: field %s is NULL", k_name[f]);
    CU(cudaMemcpyAsync(st->dev[f], host[f], st->bytes[f], cudaMemcpyHostToDevice, (cudaStream_t)stream));
  }
  CU(cudaMemcpyAsync(st->plude0, st->dev[CLOUDSC_F_PLUDE], st->bytes[CLOUDSC_F_PLUDE], cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int cloudsc_b200_download(cloudsc_b200_state *st, void *const host[CLOUDSC_NFIELDS], void *stream) {
  if (!st || !host) return fail(CLOUDSC_B200_EARG, "download: NULL argument");
  if (int rc = state_on_current_device(st, "download")) return rc;
  for (int f = 0; f < CLOUDSC_NFIELDS; ++f) {
    if (!is_out(f)) continue;
    if (!host[f]) return fail(CLOUDSC_B200_EARG, "download: field %s is NULL", k_name[f]);
    CU(cudaMemcpyAsync(host[f], st->dev[f], st->bytes[f], cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  }
  return 0;
}

int cloudsc_b200_reset_inout(cloudsc_b200_state *st, void *stream) {
  if (!st) return fail(CLOUDSC_B200_EARG, "reset_inout: NULL state");
  if (int rc = state_on_current_device(st, "reset_inout")) return rc;
  CU(cudaMemcpyAsync(st->dev[CLOUDSC_F_PLUDE], st->plude0, st->bytes[CLOUDSC_F_PLUDE], cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int cloudsc_b200_step(cloudsc_b200_state *st, double ptsphy, int first_block, int nblocks, void *stream) {
  if (!st) return fail(CLOUDSC_B200_EARG, "step: NULL state");
  if (int rc = current_device_ready()) return rc;
  if (int rc = state_on_current_device(st, "step")) return rc;
  if ((sweep_aerosol_mask() & ~st->aer) != 0)
    return fail(CLOUDSC_B200_EPARAMS, "step: aerosol switches were turned on after the state was created (it holds no aerosol fields)");
  if (nblocks <= 0) { first_block = 0; nblocks = st->nblocks; }
  if (first_block < 0 || first_block + nblocks > st->nblocks) return fail(CLOUDSC_B200_EARG, "step: block range out of bounds");
  FieldTable ft;
  for (int i = 0; i < CLOUDSC_NFIELDS; ++i) ft.p[i] = st->dev[i];
  long long end = (long long)(first_block + nblocks) * st->nproma;
  if (end > st->ngptot) end = st->ngptot;
  LaunchGeom g{first_block * st->nproma, (int)end, st->nproma, st->klev};
  const int variant = g_variant.load();
  /* a launch captured into a CUDA graph is replayed with the SAME kernel arguments: the per-launch sequence number of the
     schedules with scratch (modes 1-3) would repeat, so a capturing stream gets one CTA per tile */
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (st->sched_mode >= 1 && cudaStreamIsCapturing((cudaStream_t)stream, &cap) != cudaSuccess) { (void)cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
  if (st->has_tm && variant == VARIANT_AUTO && st->sched_mode >= 1 && cap == cudaStreamCaptureStatusNone)
    CU(launch_sweep_persist(ft, st->tm, g, ptsphy, st->precision, st->sched, (cudaStream_t)stream));
  else if (st->has_tm && variant == VARIANT_AUTO)
    CU(launch_sweep_tmap(ft, st->tm, g, ptsphy, st->precision, (cudaStream_t)stream));
  else
    CU(launch_sweep(ft, g, ptsphy, st->precision, variant, (cudaStream_t)stream));
  return 0;
}

int cloudsc_b200_state_kernel_info(cloudsc_b200_state *st, int out[8]) {
  if (!st || !out) return fail(CLOUDSC_B200_EARG, "state_kernel_info: NULL argument");
  if (int rc = state_on_current_device(st, "state_kernel_info")) return rc;
  const int variant = g_variant.load();
  if (st->has_tm && variant == VARIANT_AUTO && !sweep_aerosol_mask()) {
    int used = 0;                      /* the schedule a step() over the whole state gets */
    if (st->sched_mode >= 1) {
      LaunchGeom g{0, st->ngptot, st->nproma, st->klev};
      CU(persist_schedule_used(g, st->precision, st->sched, &used));
    }
    CU(tmap_kernel_info(st->precision, out, used));
  } else {
    CU(sweep_kernel_info(st->nproma, st->precision, variant, out));
  }
  return 0;
}

int cloudsc_b200_get_schedule(cloudsc_b200_state *st, int out[4]) {
  if (!st || !out) return fail(CLOUDSC_B200_EARG, "get_schedule: NULL argument");
  if (int rc = state_on_current_device(st, "get_schedule")) return rc;
  out[0] = st->sched_mode; out[1] = st->sched.ksplit; out[2] = st->sched.target_b; out[3] = 0;
  if (st->has_tm && g_variant.load() == VARIANT_AUTO && st->sched_mode >= 1) {
    LaunchGeom g{0, st->ngptot, st->nproma, st->klev};
    CU(persist_schedule_used(g, st->precision, st->sched, &out[3]));
  }
  return 0;
}

void *cloudsc_b200_field_ptr(cloudsc_b200_state *st, int f) {
  return (st && f >= 0 && f < CLOUDSC_NFIELDS) ? st->dev[f] : nullptr;
}

/* ---------------------------------------------------------------------------------------------- */
/* host-resident fields: chunked multi-stream pipeline                                            */
/* ---------------------------------------------------------------------------------------------- */
namespace {
double wall_now() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
/*
 * Per-device pipeline resources, kept between calls: cudaMalloc/cudaFree cost ~1 ms / ~0.4 ms each on
 * B200 (tools/ubench/pcie.cu), which for 51 fields x 3 lanes would be twice the time the copies
 * themselves need.  One arena per lane, laid out like the resident state's (arena_layout), so the lane's
 * kernel is the tensor-map kernel too.
 */
constexpr int kMaxLanes = 8;
struct Lane {
  cudaStream_t stream = nullptr;
  cudaEvent_t k0 = nullptr, k1 = nullptr;
  cudaEvent_t up = nullptr, down = nullptr;   /* ring mode: inputs uploaded / outputs downloaded (arena free again) */
  char *arena = nullptr;
  size_t bytes = 0;             /* size of `arena` */
  bool used = false;
  /* geometry the tensor maps below were encoded for */
  int tm_blocks = 0, tm_nproma = 0, tm_klev = 0, tm_precision = -1, tm_aer = -1;
  bool has_tm = false;
  TensorMaps tm;
};
struct Pipe {
  std::mutex mu;
  Lane lanes[kMaxLanes];
  cudaEvent_t w0 = nullptr, w1 = nullptr;
  /* ring mode: ONE stream per copy direction and one for the kernels, the lanes' arenas as a ring of buffers */
  cudaStream_t s_up = nullptr, s_k = nullptr, s_down = nullptr;
  void release() {
    if (s_up) cudaStreamDestroy(s_up);
    if (s_k) cudaStreamDestroy(s_k);
    if (s_down) cudaStreamDestroy(s_down);
    s_up = s_k = s_down = nullptr;
    for (int i = 0; i < kMaxLanes; ++i) {
      Lane &l = lanes[i];
      if (l.arena) cudaFree(l.arena);
      if (l.k0) cudaEventDestroy(l.k0);
      if (l.k1) cudaEventDestroy(l.k1);
      if (l.up) cudaEventDestroy(l.up);
      if (l.down) cudaEventDestroy(l.down);
      if (l.stream) cudaStreamDestroy(l.stream);
      l = Lane();
    }
    if (w0) cudaEventDestroy(w0);
    if (w1) cudaEventDestroy(w1);
    w0 = w1 = nullptr;
  }
};
Pipe g_pipes[64];

/* every lane tracks the size of ITS arena (an arena allocated for an earlier, smaller request is replaced) */
cudaError_t pipe_reserve(Pipe &p, int nlanes, size_t lane_bytes) {
  cudaError_t e;
  if (!p.w0) {
    if ((e = cudaEventCreate(&p.w0)) != cudaSuccess) return e;
    if ((e = cudaEventCreate(&p.w1)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&p.s_up, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&p.s_k, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&p.s_down, cudaStreamNonBlocking)) != cudaSuccess) return e;
  }
  for (int i = 0; i < nlanes; ++i) {
    Lane &l = p.lanes[i];
    if (!l.stream) {
      if ((e = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
      if ((e = cudaEventCreate(&l.k0)) != cudaSuccess) return e;
      if ((e = cudaEventCreate(&l.k1)) != cudaSuccess) return e;
      if ((e = cudaEventCreateWithFlags(&l.up, cudaEventDisableTiming)) != cudaSuccess) return e;
      if ((e = cudaEventCreateWithFlags(&l.down, cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    if (!l.arena || l.bytes < lane_bytes) {
      if (l.arena) { cudaFree(l.arena); l.arena = nullptr; l.bytes = 0; }
      if ((e = cudaMalloc(&l.arena, lane_bytes)) != cudaSuccess) return e;
      l.bytes = lane_bytes;
      l.has_tm = false; l.tm_blocks = 0;
      /* rows of the inputs the kernel never reads are never uploaded (CopyPlan): give them a defined value once */
      if ((e = cudaMemsetAsync(l.arena, 0, lane_bytes, l.stream)) != cudaSuccess) return e;
      if ((e = cudaStreamSynchronize(l.stream)) != cudaSuccess) return e;   /* (the ring mode uses other streams) */
    }
    l.used = false;
  }
  return cudaSuccess;
}

/*
 * What travels per field and block.  The kernel reads species 0..3 of PCLV / TENDENCY_TMP_CLD only (SURVEY App. A
 * #10, #34) and, above the first physics level ktop = ncldtop - 1, none of the rows listed below (cloudsc_column.cuh:
 * qv, the forcing terms and the convective fluxes are only used for k >= ktop; pap[ktop-1], plu[k+1] and
 * pmfu/pmfd[k+1] one level earlier / later).  It writes species 0..3 of TENDENCY_LOC_CLD; species 4 is zeroed on the
 * host, as the reference driver does (cloudsc_driver.c:200).  tests/test_gpu_parity.py::test_skipped_input_rows_are_dead
 * poisons the skipped rows and requires bit-identical results.
 */
struct CopyPlan {
  size_t in0[CLOUDSC_NFIELDS], in1[CLOUDSC_NFIELDS];     /* rows [in0, in1) of a block are uploaded */
  size_t out0[CLOUDSC_NFIELDS], out1[CLOUDSC_NFIELDS];   /* rows [out0, out1) are downloaded */
};
CopyPlan make_copy_plan(int klev, int ncldtop, int aer) {
  CopyPlan p;
  /* $CLOUDSC_B200_E2E_TRIM (measurement knob, tools/gpu_e2e_sweep.py): 2 = default, 1 = species only, 0 = whole fields */
  const char *env = getenv("CLOUDSC_B200_E2E_TRIM");
  const int trim = env && *env ? atoi(env) : 2;
  const size_t ktop = (size_t)(trim >= 2 && ncldtop > 1 ? ncldtop - 1 : 0);
  for (int f = 0; f < CLOUDSC_NFIELDS; ++f) {
    p.in0[f] = p.in1[f] = p.out0[f] = p.out1[f] = 0;
    if (is_in(f, aer)) p.in1[f] = rows_of(f, klev);
    if (is_out(f)) p.out1[f] = rows_of(f, klev);
  }
  if (trim >= 1) {
    p.in1[CLOUDSC_F_PCLV] = p.in1[CLOUDSC_F_TENDENCY_TMP_CLD] = (size_t)4 * klev;
    p.out1[CLOUDSC_F_TENDENCY_LOC_CLD] = (size_t)4 * klev;
  }
  static const int from_ktop[] = {CLOUDSC_F_PQ, CLOUDSC_F_TENDENCY_TMP_Q, CLOUDSC_F_PSUPSAT, CLOUDSC_F_PSNDE, CLOUDSC_F_PHRSW,
                                  CLOUDSC_F_PHRLW, CLOUDSC_F_PVERVEL, CLOUDSC_F_PMFU, CLOUDSC_F_PMFD};
  for (int f : from_ktop) p.in0[f] = ktop < (size_t)klev ? ktop : (size_t)klev;
  p.in0[CLOUDSC_F_PAP] = ktop > 0 ? ktop - 1 : 0;
  p.in0[CLOUDSC_F_PLU] = trim >= 2 ? (ktop + 1 < (size_t)klev ? ktop + 1 : (size_t)klev) : 0;
  for (int f = 0; f < CLOUDSC_NFIELDS; ++f) {
    if (is_aer_in(f, aer)) p.in0[f] = ktop < (size_t)klev ? ktop : (size_t)klev;
    if (p.in0[f] > p.in1[f]) p.in0[f] = p.in1[f];
  }
  return p;
}

/* rows [r0, r1) of `nb` consecutive blocks of one field, either direction; returns the bytes moved */
cudaError_t copy_rows(char *dst, const char *src, size_t blk_bytes, size_t row_bytes, size_t r0, size_t r1, int nb,
                      cudaMemcpyKind kind, cudaStream_t stream, size_t *moved) {
  if (r1 <= r0 || nb <= 0) return cudaSuccess;
  const size_t width = (r1 - r0) * row_bytes;
  *moved += width * (size_t)nb;
  if (width == blk_bytes) return cudaMemcpyAsync(dst, src, blk_bytes * (size_t)nb, kind, stream);
  return cudaMemcpy2DAsync(dst + r0 * row_bytes, blk_bytes, src + r0 * row_bytes, blk_bytes, width, (size_t)nb, kind, stream);
}
}  // namespace

int cloudsc_b200_release_cache(void) {
  for (auto &p : g_pipes) {
    std::lock_guard<std::mutex> lk(p.mu);
    p.release();
  }
  return 0;
}

int cloudsc_b200_run_host_ex(void *const host[CLOUDSC_NFIELDS], int ngptot, int nproma, int klev, double ptsphy,
                             int precision, int chunk_blocks, int nstreams, double stats[4]) {
  if (int rc = check_geom(ngptot, nproma, klev)) return rc;
  if (precision != CLOUDSC_B200_FP64 && precision != CLOUDSC_B200_FP32) return fail(CLOUDSC_B200_EARG, "bad precision %d", precision);
  if (int rc = current_device_ready()) return rc;
  const int aer = sweep_aerosol_mask();
  if (int rc = check_fields(host, aer)) return rc;
  const int nblocks = (ngptot + nproma - 1) / nproma;
  if (nstreams <= 0) nstreams = 4;
  if (nstreams > kMaxLanes) nstreams = kMaxLanes;
  /* chunk schedule.  Default: ~5 chunks per lane, so that pipeline fill (first upload) and drain (last download)
     stay small against the transfer time.  The kernel is ~2.5 % of a chunk's copy time: chunks are sized for overlap,
     not for kernel occupancy; the floor only avoids degenerate launches.  Measured on one box, alternating
     (tools/gpu_e2e_align.py, profiles/r2_e2e_align.txt): 20 equal chunks of 64 blocks 108.5 ms, 23 of 56 110.0 ms, 56 with
     half- and quarter-size chunks at both ends ($CLOUDSC_B200_E2E_RAMP=1, the earlier default) 112.0 ms -- every chunk
     costs its 50 copies, and the small ones at the ends cost more than the fill / drain time they save. */
  std::vector<int> sizes;
  /* Every chunk starts on a block index that puts the host side of all its copies on a 4 KB boundary (one NPROMA row of
     doubles is 1 KB at NPROMA 128: multiples of 4 blocks).  Copies that start inside a page cost 6 % of the PCIe rate
     (tools/ubench/pcie.cu --pattern 2 --knobs 2, profiles/r2_pcie_pattern.txt).  $CLOUDSC_B200_E2E_ALIGN=0 turns it off. */
  int ab = 1;
  {
    const char *al_env = getenv("CLOUDSC_B200_E2E_ALIGN");
    if (!(al_env && *al_env == '0')) {
      size_t rowbytes = (size_t)nproma * (precision == CLOUDSC_B200_FP32 ? 4 : 8), gcd = 4096, t = rowbytes % 4096;
      while (t) { const size_t r = gcd % t; gcd = t; t = r; }
      ab = (int)(4096 / gcd);
      if (ab > 64) ab = 1;   /* odd NPROMA: no sensible alignment */
    }
  }
  auto round_ab = [&](int n) { return n <= ab ? ab : (n + ab / 2) / ab * ab; };
  if (chunk_blocks <= 0) {
    chunk_blocks = (nblocks + nstreams * 5 - 1) / (nstreams * 5);
    const int min_blocks = (4096 + nproma - 1) / nproma;
    if (chunk_blocks < min_blocks) chunk_blocks = min_blocks;
    chunk_blocks = round_ab(chunk_blocks);
    if (chunk_blocks > nblocks) chunk_blocks = nblocks;
    int left = nblocks;
    /* quarter- and half-size chunks at both ends: the first upload and the last download overlap with nothing
       (measured 3.7 + 2.3 ms of a 110 ms call with full-size end chunks, profiles/r2_e2e_trace.txt) */
    const int q = round_ab(chunk_blocks / 4), h = round_ab(chunk_blocks / 2);
    const int ramp_floor = (1024 + nproma - 1) / nproma;
    const char *ramp_env = getenv("CLOUDSC_B200_E2E_RAMP");
    if (ramp_env && *ramp_env == '1' && q >= ramp_floor && left > 4 * chunk_blocks) {
      sizes.push_back(q); sizes.push_back(h);
      left -= q + h;
      const int tail = q + h;
      while (left - tail >= chunk_blocks) { sizes.push_back(chunk_blocks); left -= chunk_blocks; }
      if (left > tail) { sizes.push_back(left - tail); left = tail; }
      sizes.push_back(left - q);
      sizes.push_back(q);
    } else {
      while (left > 0) { const int n = left < chunk_blocks ? left : chunk_blocks; sizes.push_back(n); left -= n; }
    }
  } else {
    if (chunk_blocks > nblocks) chunk_blocks = nblocks;
    for (int left = nblocks; left > 0; left -= chunk_blocks) sizes.push_back(left < chunk_blocks ? left : chunk_blocks);
  }
  const int nchunks = (int)sizes.size();
  if (nstreams > nchunks) nstreams = nchunks;

  /* arena layout of one lane, per-field block and row sizes, what is copied */
  const ArenaLayout lay = arena_layout(chunk_blocks, nproma, klev, precision, aer);
  size_t blk[CLOUDSC_NFIELDS], rowb[CLOUDSC_NFIELDS];
  for (int f = 0; f < CLOUDSC_NFIELDS; ++f) {
    rowb[f] = (size_t)nproma * elem_size(f, precision);
    blk[f] = rows_of(f, klev) * rowb[f];
  }
  const CopyPlan plan = make_copy_plan(klev, sweep_ncldtop(), aer);
  const int valid_last = ngptot - (nblocks - 1) * nproma;   /* columns of the last block that exist for the caller */
  const bool ragged = valid_last < nproma;

  int dev = 0;
  CU(cudaGetDevice(&dev));
  Pipe &pipe = g_pipes[dev & 63];
  std::lock_guard<std::mutex> lk(pipe.mu);
  {
    cudaError_t e = pipe_reserve(pipe, nstreams, lay.total);
    if (e != cudaSuccess) { pipe.release(); return cuda_fail(e, "run_host: pipeline resources"); }
  }
  const int variant = g_variant.load();
  /* $CLOUDSC_B200_E2E_COPIES_ONLY=1 (tools/gpu_e2e_sweep.py): skip the kernel launches, to time the transfers alone */
  const char *co_env = getenv("CLOUDSC_B200_E2E_COPIES_ONLY");
  const bool copies_only = co_env && *co_env == '1';
  for (int i = 0; i < nstreams; ++i) {
    Lane &l = pipe.lanes[i];
    if (l.tm_blocks != chunk_blocks || l.tm_nproma != nproma || l.tm_klev != klev || l.tm_precision != precision || l.tm_aer != aer) {
      void *d[CLOUDSC_NFIELDS];
      for (int f = 0; f < CLOUDSC_NFIELDS; ++f) d[f] = is_live(f, aer) ? l.arena + lay.off[f] : nullptr;
      l.has_tm = !aer && build_tensor_maps(d, chunk_blocks, nproma, klev, precision, &l.tm);
      l.tm_blocks = chunk_blocks; l.tm_nproma = nproma; l.tm_klev = klev; l.tm_precision = precision; l.tm_aer = aer;
    }
  }

  double kernel_ms = 0.0;
  size_t h2d = 0, d2h = 0;
  /* on error: drain what is in flight, then report (the cached resources stay valid) */
#define CUX(call)                                                    \
  do {                                                               \
    cudaError_t e_ = (call);                                         \
    if (e_ != cudaSuccess) {                                         \
      for (int i_ = 0; i_ < nstreams; ++i_) cudaStreamSynchronize(pipe.lanes[i_].stream); \
      cudaStreamSynchronize(pipe.s_up); cudaStreamSynchronize(pipe.s_k); cudaStreamSynchronize(pipe.s_down); \
      return cuda_fail(e_, #call);                                   \
    }                                                                \
  } while (0)

  CUX(cudaEventRecord(pipe.w0, 0));
  /* the reference driver zeroes pcovptot and all of tend_loc_cld of a block before the call (cloudsc_driver.c:199-200);
     the kernel writes every valid element of both, so only the padding columns of a ragged last block still need it */
  if (ragged) {
    const size_t es = elem_size(CLOUDSC_F_PCOVPTOT, precision);
    char *cov = (char *)host[CLOUDSC_F_PCOVPTOT] + (size_t)(nblocks - 1) * blk[CLOUDSC_F_PCOVPTOT];
    char *cld = (char *)host[CLOUDSC_F_TENDENCY_LOC_CLD] + (size_t)(nblocks - 1) * blk[CLOUDSC_F_TENDENCY_LOC_CLD];
    for (size_t r = 0; r < (size_t)klev; ++r) memset(cov + r * rowb[CLOUDSC_F_PCOVPTOT] + (size_t)valid_last * es, 0, (size_t)(nproma - valid_last) * es);
    for (size_t r = 0; r < (size_t)4 * klev; ++r) memset(cld + r * rowb[CLOUDSC_F_TENDENCY_LOC_CLD] + (size_t)valid_last * es, 0, (size_t)(nproma - valid_last) * es);
  }
  /* Stream layout.  ring (default): ONE upload stream, ONE kernel stream, ONE download stream; the lanes' arenas form
     a ring of buffers handed from stream to stream by events -- each copy engine is fed by a single in-order queue,
     as in the two-stream copy micro-benchmark (tools/ubench/pcie.cu).  lanes ($CLOUDSC_B200_E2E_PIPE=lanes, the
     round-1 layout): every lane has its own stream carrying upload, kernel and download of its chunks. */
  const char *pipe_env = getenv("CLOUDSC_B200_E2E_PIPE");
  const bool ring = !(pipe_env && strcmp(pipe_env, "lanes") == 0);
  /* $CLOUDSC_B200_E2E_TRACE=1: time line of the call on stderr (per chunk: upload / kernel / download begin and end on
     the device, host-side phases), for tools/gpu_e2e_sweep.py --trace */
  const char *tr_env = getenv("CLOUDSC_B200_E2E_TRACE");
  const bool trace = tr_env && *tr_env == '1' && nchunks <= 96;
  std::vector<cudaEvent_t> tev;
  if (trace) {
    tev.resize((size_t)nchunks * 6);
    for (auto &e : tev) cudaEventCreate(&e);
  }
  const double host_t0 = wall_now();
  double host_enq = 0.0, host_memset = 0.0;
  int b0 = 0;
  for (int c = 0; c < nchunks; ++c) {
    Lane &l = pipe.lanes[c % nstreams];
    const int nb = sizes[c];
    cudaStream_t s_up = ring ? pipe.s_up : l.stream, s_k = ring ? pipe.s_k : l.stream, s_down = ring ? pipe.s_down : l.stream;
    if (l.used) {  /* harvest the kernel time of the chunk that used this lane before */
      CUX(cudaEventSynchronize(l.k1));
      float ms = 0.f;
      CUX(cudaEventElapsedTime(&ms, l.k0, l.k1));
      kernel_ms += ms;
      if (ring) CUX(cudaStreamWaitEvent(s_up, l.down, 0));   /* the arena is free once its outputs are on the host */
    }
    FieldTable ft;
    for (int f = 0; f < CLOUDSC_NFIELDS; ++f) ft.p[f] = is_live(f, aer) ? l.arena + lay.off[f] : nullptr;
    if (trace) cudaEventRecord(tev[(size_t)c * 6 + 0], s_up);
    for (int f = 0; f < CLOUDSC_NFIELDS; ++f) {
      if (plan.in1[f] <= plan.in0[f]) continue;
      CUX(copy_rows((char *)ft.p[f], (const char *)host[f] + (size_t)b0 * blk[f], blk[f], rowb[f], plan.in0[f], plan.in1[f], nb,
                    cudaMemcpyHostToDevice, s_up, &h2d));
    }
    if (trace) cudaEventRecord(tev[(size_t)c * 6 + 1], s_up);
    if (ring) { CUX(cudaEventRecord(l.up, s_up)); CUX(cudaStreamWaitEvent(s_k, l.up, 0)); }
    long long cend = (long long)(b0 + nb) * nproma;
    if (cend > ngptot) cend = ngptot;
    LaunchGeom g{0, (int)(cend - (long long)b0 * nproma), nproma, klev};
    CUX(cudaEventRecord(l.k0, s_k));
    if (copies_only) { /* measurement knob: the copy pipeline alone (outputs are then meaningless) */ }
    else if (l.has_tm && variant == VARIANT_AUTO) CUX(launch_sweep_tmap(ft, l.tm, g, ptsphy, precision, s_k));
    else CUX(launch_sweep(ft, g, ptsphy, precision, variant, s_k));
    CUX(cudaEventRecord(l.k1, s_k));
    if (ring) CUX(cudaStreamWaitEvent(s_down, l.k1, 0));
    if (trace) { cudaEventRecord(tev[(size_t)c * 6 + 2], s_k); cudaEventRecord(tev[(size_t)c * 6 + 3], s_down); }
    /* downloads: whole blocks, except that of a ragged last block only the columns the caller's run covers come
       back (the reference leaves the padding columns of its outputs untouched) */
    const bool has_ragged = ragged && b0 + nb == nblocks;
    const int nfull = has_ragged ? nb - 1 : nb;
    for (int f = 0; f < CLOUDSC_NFIELDS; ++f) {
      if (plan.out1[f] <= plan.out0[f]) continue;
      CUX(copy_rows((char *)host[f] + (size_t)b0 * blk[f], (const char *)ft.p[f], blk[f], rowb[f], plan.out0[f], plan.out1[f], nfull,
                    cudaMemcpyDeviceToHost, s_down, &d2h));
      if (has_ragged) {
        const size_t es = elem_size(f, precision), w = (size_t)valid_last * es, rows = plan.out1[f] - plan.out0[f];
        CUX(cudaMemcpy2DAsync((char *)host[f] + (size_t)(nblocks - 1) * blk[f] + plan.out0[f] * rowb[f], rowb[f],
                              (const char *)ft.p[f] + (size_t)nfull * blk[f] + plan.out0[f] * rowb[f], rowb[f], w, rows,
                              cudaMemcpyDeviceToHost, s_down));
        d2h += w * rows;
      }
    }
    if (trace) cudaEventRecord(tev[(size_t)c * 6 + 4], s_down);
    if (ring) CUX(cudaEventRecord(l.down, s_down));
    l.used = true;
    b0 += nb;
  }
  host_enq = wall_now() - host_t0;
  /* species 4 (vapour) of TENDENCY_LOC_CLD is never written by the kernel and never downloaded: it holds the
     zeros of cloudsc_driver.c:200.  Written here, by the otherwise idle calling thread, while the copies run. */
  {
    const size_t es = elem_size(CLOUDSC_F_TENDENCY_LOC_CLD, precision);
    const size_t sp = (size_t)klev * nproma * es;
    char *cld = (char *)host[CLOUDSC_F_TENDENCY_LOC_CLD];
    for (int b = 0; b < nblocks; ++b) memset(cld + (size_t)b * blk[CLOUDSC_F_TENDENCY_LOC_CLD] + 4 * sp, 0, sp);
  }
  host_memset = wall_now() - host_t0;
  if (ring) { CUX(cudaStreamSynchronize(pipe.s_down)); CUX(cudaStreamSynchronize(pipe.s_k)); CUX(cudaStreamSynchronize(pipe.s_up)); }
  for (int i = 0; i < nstreams; ++i) {
    Lane &l = pipe.lanes[i];
    CUX(cudaStreamSynchronize(l.stream));
    if (l.used) {
      float ms = 0.f;
      CUX(cudaEventElapsedTime(&ms, l.k0, l.k1));
      kernel_ms += ms;
    }
  }
  CUX(cudaEventRecord(pipe.w1, 0));
  CUX(cudaEventSynchronize(pipe.w1));
  float wall = 0.f;
  CUX(cudaEventElapsedTime(&wall, pipe.w0, pipe.w1));
  if (stats) { stats[0] = (double)h2d; stats[1] = (double)d2h; stats[2] = kernel_ms; stats[3] = wall; }
  if (trace) {
    fprintf(stderr, "run_host trace: %d chunks, %d buffers, %s; host: enqueue done %.2f ms, species-5 zeroing done %.2f ms, all done %.2f ms\n",
            nchunks, nstreams, ring ? "ring" : "lanes", 1e3 * host_enq, 1e3 * host_memset, 1e3 * (wall_now() - host_t0));
    fprintf(stderr, "  chunk blocks |  upload begin..end  | kernel end | download begin..end   [ms since the first upload began]\n");
    for (int c = 0; c < nchunks; ++c) {
      float t[5] = {0, 0, 0, 0, 0};
      for (int j = 0; j < 5; ++j) cudaEventElapsedTime(&t[j], tev[0], tev[(size_t)c * 6 + j]);
      fprintf(stderr, "  %5d %6d | %8.2f .. %8.2f | %8.2f | %8.2f .. %8.2f\n", c, sizes[c], t[0], t[1], t[2], t[3], t[4]);
    }
    for (auto &e : tev) cudaEventDestroy(e);
  }
#undef CUX
  return 0;
}

int cloudsc_b200_run_host(void *const host[CLOUDSC_NFIELDS], int ngptot, int nproma, int klev, double ptsphy,
                          int chunk_blocks, int nstreams, double stats[4]) {
  return cloudsc_b200_run_host_ex(host, ngptot, nproma, klev, ptsphy, CLOUDSC_B200_FP64, chunk_blocks, nstreams, stats);
}

/* ---------------------------------------------------------------------------------------------- */
/* drop-in cloudsc_c(): one NPROMA block, host pointers, reference globals                        */
/* ---------------------------------------------------------------------------------------------- */
namespace {
struct ShimBuf {
  void *dev[CLOUDSC_NFIELDS] = {nullptr};
  int klon = 0, klev = 0, device = -1, aer = -1;
  cudaStream_t stream = nullptr;
  bool seen = false;
  struct TECLDP seen_val;
  struct cloudsc_consts seen_k;
  void release() {
    if (device >= 0) {
      int cur = -1;
      cudaGetDevice(&cur);
      if (cur != device) cudaSetDevice(device);
      for (int f = 0; f < CLOUDSC_NFIELDS; ++f)
        if (dev[f]) { cudaFree(dev[f]); dev[f] = nullptr; }
      if (stream) { cudaStreamDestroy(stream); stream = nullptr; }
      if (cur >= 0 && cur != device) cudaSetDevice(cur);
    }
    klon = klev = 0; device = -1; aer = -1; seen = false;
  }
  /* thread exit (an OpenMP worker being retired, the end of the process): give the staging buffers back.  At process
     exit the CUDA context may already be gone; cudaFree then fails harmlessly. */
  ~ShimBuf() { release(); }
};
thread_local ShimBuf t_shim;

/* member-wise comparison: struct TECLDP has padding after its 17 ints and after nbeta, and the reference allocates it
 * with malloc, so memcmp would see differences that are not there */
bool same_params(const struct TECLDP &a, const struct TECLDP &b) {
#define CLOUDSC_X(n) if (!(a.n == b.n)) return false;
  CLOUDSC_TECLDP_REALS(CLOUDSC_X)
  CLOUDSC_TECLDP_INTS(CLOUDSC_X)
#undef CLOUDSC_X
  return a.nshapep == b.nshapep && a.nshapeq == b.nshapeq && a.nbeta == b.nbeta;
}
bool same_consts(const struct cloudsc_consts &a, const struct cloudsc_consts &b) {
#define CLOUDSC_X(n) if (!(a.n == b.n)) return false;
  CLOUDSC_KERNEL_CONSTS(CLOUDSC_X)
#undef CLOUDSC_X
  return true;
}
}  // namespace

/* frees the calling thread's cloudsc_c() staging buffers and stream (they are also freed at thread exit) */
int cloudsc_b200_release_thread(void) {
  t_shim.release();
  return 0;
}

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
  double *v_pfplsl, double *v_pfplsn, double *v_pfhpsl, double *v_pfhpsn) {
  void *host[CLOUDSC_NFIELDS] = {
    v_pt, v_pq, v_tendency_cml_t, v_tendency_cml_q, v_tendency_cml_a, v_tendency_cml_cld,
    v_tendency_tmp_t, v_tendency_tmp_q, v_tendency_tmp_a, v_tendency_tmp_cld,
    v_tendency_loc_t, v_tendency_loc_q, v_tendency_loc_a, v_tendency_loc_cld,
    v_pvfa, v_pvfl, v_pvfi, v_pdyna, v_pdynl, v_pdyni, v_phrsw, v_phrlw, v_pvervel, v_pap, v_paph, v_plsm,
    v_ktype, v_plu, v_plude, v_psnde, v_pmfu, v_pmfd, v_pa, v_pclv, v_psupsat, v_plcrit_aer, v_picrit_aer,
    v_pre_ice, v_pccn, v_pnice, v_pcovptot, v_prainfrac_toprfz, v_pfsqlf, v_pfsqif, v_pfcqnng, v_pfcqlng,
    v_pfsqrf, v_pfsqsf, v_pfcqrng, v_pfcqsng, v_pfsqltur, v_pfsqitur, v_pfplsl, v_pfplsn, v_pfhpsl, v_pfhpsn};
  if (kidia < 1 || kfdia > klon || kfdia < kidia) return fail(CLOUDSC_B200_EARG, "cloudsc_c: bad column range %d..%d of %d", kidia, kfdia, klon);
  if (int rc = check_geom(klon, klon, klev)) return rc;
  if (!yrecldp) return fail(CLOUDSC_B200_EARG, "cloudsc_c: yrecldp is NULL");

  ShimBuf &s = t_shim;
  int dev = 0;
  CU(cudaGetDevice(&dev));
  /* re-snapshot the globals whenever they changed since this thread's last call (set_params() itself skips the
     upload when the device already holds the same values) */
  struct cloudsc_consts k;
#define CLOUDSC_X(n) k.n = n;
  CLOUDSC_KERNEL_CONSTS(CLOUDSC_X)
#undef CLOUDSC_X
  bool ready;
  { std::lock_guard<std::mutex> lk(g_mu); ready = dev >= 0 && dev < 64 && g_params_set[dev]; }
  if (!ready || !s.seen || s.device != dev || !same_params(s.seen_val, *yrecldp) || !same_consts(s.seen_k, k)) {
    if (int rc = cloudsc_b200_set_params(yrecldp, &k)) return rc;
    s.seen_val = *yrecldp;
    s.seen_k = k;
    s.seen = true;
  }
  const int aer = aerosol_mask(yrecldp);
  if (int rc = check_fields(host, aer)) return rc;
  if (s.device != dev || s.klon != klon || s.klev != klev || s.aer != aer) {
    s.release();
    s.device = dev;
    for (int f = 0; f < CLOUDSC_NFIELDS; ++f) {
      if (!is_live(f, aer)) continue;
      CU(cudaMalloc(&s.dev[f], rows_of(f, klev) * klon * elem_size(f, CLOUDSC_B200_FP64)));
    }
    CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    s.klon = klon; s.klev = klev; s.aer = aer;
    s.seen_val = *yrecldp; s.seen_k = k; s.seen = true;
  }
  const int ncols = kfdia - kidia + 1;
  for (int f = 0; f < CLOUDSC_NFIELDS; ++f) {
    if (!is_in(f, aer)) continue;
    const size_t es = elem_size(f, CLOUDSC_B200_FP64);
    size_t rows = rows_of(f, klev);
    if (f == CLOUDSC_F_PCLV || f == CLOUDSC_F_TENDENCY_TMP_CLD) rows = (size_t)4 * klev;   /* species 5 is never read */
    CU(cudaMemcpy2DAsync((char *)s.dev[f] + (size_t)(kidia - 1) * es, (size_t)klon * es,
                         (const char *)host[f] + (size_t)(kidia - 1) * es, (size_t)klon * es,
                         (size_t)ncols * es, rows, cudaMemcpyHostToDevice, s.stream));
  }
  FieldTable ft;
  for (int i = 0; i < CLOUDSC_NFIELDS; ++i) ft.p[i] = s.dev[i];
  LaunchGeom g{kidia - 1, kfdia, klon, klev};
  CU(launch_sweep(ft, g, ptsphy, CLOUDSC_B200_FP64, g_variant.load(), s.stream));
  for (int f = 0; f < CLOUDSC_NFIELDS; ++f) {
    if (!is_out(f)) continue;
    const size_t es = elem_size(f, CLOUDSC_B200_FP64);
    size_t rows = rows_of(f, klev);
    char *h = (char *)host[f];
    const char *d = (const char *)s.dev[f];
    if (f == CLOUDSC_F_TENDENCY_LOC_CLD) rows = (size_t)4 * klev;  /* species 5 is the caller's (cloudsc_driver.c:200) */
    CU(cudaMemcpy2DAsync(h + (size_t)(kidia - 1) * es, (size_t)klon * es, d + (size_t)(kidia - 1) * es, (size_t)klon * es,
                         (size_t)ncols * es, rows, cudaMemcpyDeviceToHost, s.stream));
  }
  CU(cudaStreamSynchronize(s.stream));
  return 0;
}

}  /* extern "C" */
