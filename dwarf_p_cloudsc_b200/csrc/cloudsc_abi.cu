/*
 * cloudsc_abi.cu -- the extern "C" boundary of libcloudsc_b200.so (declared in include/cloudsc_b200.h).
 *
 * Host-side plumbing only: parameter snapshot into __constant__ memory, device-resident launch,
 * the pinned/chunked/multi-stream host pipeline, the resident-state handle API and the
 * binary-compatible cloudsc_c() shim.  There is NO host arithmetic path in this library: every
 * entry point either runs the sm_100a kernels of cloudsc_kernels.cu or returns an error.
 */
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

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

inline bool is_in(int f) { return k_use[f] == 'R' || k_use[f] == 'B'; }
inline bool is_out(int f) { return k_use[f] == 'W' || k_use[f] == 'B'; }
inline bool is_live(int f) { return is_in(f) || is_out(f); }
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
int g_variant = 0;

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

int check_fields(void *const f[CLOUDSC_NFIELDS]) {
  if (!f) return fail(CLOUDSC_B200_EARG, "field table is NULL");
  for (int i = 0; i < CLOUDSC_NFIELDS; ++i)
    if (is_live(i) && !f[i]) return fail(CLOUDSC_B200_EARG, "field %s is NULL", k_name[i]);
  return 0;
}

}  // namespace

/* ============================================================================================== */
extern "C" {

const char *cloudsc_b200_version(void) { return "cloudsc_b200 0.1 (sm_100a)"; }
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

int cloudsc_b200_host_free(void *p) {
  if (p) CU(cudaFreeHost(p));
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

int cloudsc_b200_set_params(const struct TECLDP *p, const struct cloudsc_consts *k) {
  if (!p || !k) return fail(CLOUDSC_B200_EARG, "set_params: NULL argument");
  if (const char *why = check_params(p)) return fail(CLOUDSC_B200_EPARAMS, "unsupported parameter set: %s", why);
  CloudscConst<double> c64;
  CloudscConst<float> c32;
  fill_consts<double>(p, k, &c64);
  fill_consts<float>(p, k, &c32);
  int dev = 0;
  CU(cudaGetDevice(&dev));
  CU(upload_consts(c64, c32));
  std::lock_guard<std::mutex> lk(g_mu);
  if (dev >= 0 && dev < 64) g_params_set[dev] = true;
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
  if (variant < 0 || variant >= VARIANT_COUNT) return fail(CLOUDSC_B200_EARG, "unknown kernel variant %d", variant);
  g_variant = variant;
  return 0;
}
int cloudsc_b200_get_variant(void) { return g_variant; }

int cloudsc_b200_kernel_info(int nproma, int precision, int out[8]) {
  if (!out) return fail(CLOUDSC_B200_EARG, "kernel_info: out is NULL");
  CU(sweep_kernel_info(nproma, precision, g_variant, out));
  return 0;
}

int cloudsc_b200_launch(void *const dev_fields[CLOUDSC_NFIELDS], int ngptot, int nproma, int klev, double ptsphy,
                        int precision, void *stream) {
  if (int rc = check_geom(ngptot, nproma, klev)) return rc;
  if (int rc = check_fields(dev_fields)) return rc;
  if (precision != CLOUDSC_B200_FP64 && precision != CLOUDSC_B200_FP32) return fail(CLOUDSC_B200_EARG, "bad precision %d", precision);
  if (int rc = current_device_ready()) return rc;
  FieldTable ft;
  for (int i = 0; i < CLOUDSC_NFIELDS; ++i) ft.p[i] = dev_fields[i];
  LaunchGeom g{0, ngptot, nproma, klev};
  CU(launch_sweep(ft, g, ptsphy, precision, g_variant, (cudaStream_t)stream));
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
  void *dev[CLOUDSC_NFIELDS];   /* pointers into `arena` */
  void *plude0;                 /* pristine copy of the in/out field */
  size_t bytes[CLOUDSC_NFIELDS];
  char *arena;                  /* ONE allocation: fields of a kind at a constant spacing (see TensorMaps) */
  TensorMaps tm;
  bool has_tm;                  /* tensor maps encoded: step() uses the 4-requests-per-level TMA kernel */
};

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
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
inline size_t up256(size_t n) { return (n + 255) / 256 * 256; }
}  // namespace

int cloudsc_b200_create(int ngptot, int nproma, int klev, int precision, cloudsc_b200_state **out) {
  if (!out) return fail(CLOUDSC_B200_EARG, "create: out is NULL");
  *out = nullptr;
  if (int rc = check_geom(ngptot, nproma, klev)) return rc;
  if (precision != CLOUDSC_B200_FP64 && precision != CLOUDSC_B200_FP32) return fail(CLOUDSC_B200_EARG, "bad precision %d", precision);
  cloudsc_b200_state *st = (cloudsc_b200_state *)calloc(1, sizeof(*st));
  if (!st) return fail(CLOUDSC_B200_ENOMEM, "out of host memory");
  st->ngptot = ngptot; st->nproma = nproma; st->klev = klev; st->precision = precision;
  st->nblocks = (ngptot + nproma - 1) / nproma;
  cudaGetDevice(&st->device);

  /* arena layout: the three tensor-map groups first, each with a constant field spacing, then the rest */
  size_t off[CLOUDSC_NFIELDS];
  bool placed[CLOUDSC_NFIELDS] = {false};
  size_t total = 0;
  for (int f = 0; f < CLOUDSC_NFIELDS; ++f) {
    off[f] = 0;
    st->bytes[f] = is_live(f) ? (size_t)st->nblocks * rows_of(f, klev) * nproma * elem_size(f, precision) : 0;
  }
  auto place = [&](int f) { off[f] = total; total += up256(st->bytes[f]); placed[f] = true; };
  for (int f : kArenaGroupA) place(f);
  for (int f : kArenaGroupB) place(f);
  for (int f : kArenaGroupC) place(f);
  place(CLOUDSC_F_PAPH);
  for (int f = 0; f < CLOUDSC_NFIELDS; ++f)
    if (is_live(f) && !placed[f]) place(f);
  cudaError_t e = cudaMalloc(&st->arena, total);
  if (e != cudaSuccess) { cloudsc_b200_destroy(st); return cuda_fail(e, "cudaMalloc(arena)"); }
  for (int f = 0; f < CLOUDSC_NFIELDS; ++f) st->dev[f] = is_live(f) ? st->arena + off[f] : nullptr;
  e = cudaMalloc(&st->plude0, st->bytes[CLOUDSC_F_PLUDE]);
  if (e != cudaSuccess) { cloudsc_b200_destroy(st); return cuda_fail(e, "cudaMalloc(plude0)"); }

  /* tensor maps (whole 32-column rows: NPROMA must be a multiple of 32) */
  st->has_tm = false;
  if (nproma % 32 == 0) {
    const cuuint64_t np = (cuuint64_t)nproma, kl = (cuuint64_t)klev, nb = (cuuint64_t)st->nblocks;
    const cuuint64_t rowb = np * elem_size(CLOUDSC_F_PT, precision), SL = up256(st->bytes[CLOUDSC_F_PT]), SC = up256(st->bytes[CLOUDSC_F_PCLV]);
    const cuuint64_t dA[4] = {np, kl, nb, 15}, sA[3] = {rowb, kl * rowb, SL};
    const cuuint32_t bA[4] = {32, 1, 1, 15};
    const cuuint64_t dB[4] = {np, kl, nb, 3};
    const cuuint32_t bB[4] = {32, 1, 1, 3};
    const cuuint64_t dC[5] = {np, kl, 5, nb, 2}, sC[4] = {rowb, kl * rowb, 5 * kl * rowb, SC};
    const cuuint32_t bC[5] = {32, 1, 4, 1, 2};
    const cuuint64_t dH[3] = {np, kl + 1, nb}, sH[2] = {rowb, (kl + 1) * rowb};
    const cuuint32_t bH[3] = {32, 1, 1};
    st->has_tm = encode_map(&st->tm.a, precision, st->dev[kArenaGroupA[0]], 4, dA, sA, bA) &&
                 encode_map(&st->tm.b, precision, st->dev[kArenaGroupB[0]], 4, dB, sA, bB) &&
                 encode_map(&st->tm.c, precision, st->dev[kArenaGroupC[0]], 5, dC, sC, bC) &&
                 encode_map(&st->tm.h, precision, st->dev[CLOUDSC_F_PAPH], 3, dH, sH, bH);
  }
  *out = st;
  return 0;
}

int cloudsc_b200_destroy(cloudsc_b200_state *st) {
  if (!st) return 0;
  if (st->arena) cudaFree(st->arena);
  if (st->plude0) cudaFree(st->plude0);
  free(st);
  return 0;
}

int cloudsc_b200_upload(cloudsc_b200_state *st, void *const host[CLOUDSC_NFIELDS], void *stream) {
  if (!st || !host) return fail(CLOUDSC_B200_EARG, "upload: NULL argument");
  for (int f = 0; f < CLOUDSC_NFIELDS; ++f) {
    if (!is_in(f)) continue;
    if (!host[f]) return fail(CLOUDSC_B200_EARG, "upload: field %s is NULL", k_name[f]);
    CU(cudaMemcpyAsync(st->dev[f], host[f], st->bytes[f], cudaMemcpyHostToDevice, (cudaStream_t)stream));
  }
  CU(cudaMemcpyAsync(st->plude0, st->dev[CLOUDSC_F_PLUDE], st->bytes[CLOUDSC_F_PLUDE], cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int cloudsc_b200_download(cloudsc_b200_state *st, void *const host[CLOUDSC_NFIELDS], void *stream) {
  if (!st || !host) return fail(CLOUDSC_B200_EARG, "download: NULL argument");
  for (int f = 0; f < CLOUDSC_NFIELDS; ++f) {
    if (!is_out(f)) continue;
    if (!host[f]) return fail(CLOUDSC_B200_EARG, "download: field %s is NULL", k_name[f]);
    CU(cudaMemcpyAsync(host[f], st->dev[f], st->bytes[f], cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  }
  return 0;
}

int cloudsc_b200_reset_inout(cloudsc_b200_state *st, void *stream) {
  if (!st) return fail(CLOUDSC_B200_EARG, "reset_inout: NULL state");
  CU(cudaMemcpyAsync(st->dev[CLOUDSC_F_PLUDE], st->plude0, st->bytes[CLOUDSC_F_PLUDE], cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int cloudsc_b200_step(cloudsc_b200_state *st, double ptsphy, int first_block, int nblocks, void *stream) {
  if (!st) return fail(CLOUDSC_B200_EARG, "step: NULL state");
  if (int rc = current_device_ready()) return rc;
  if (nblocks <= 0) { first_block = 0; nblocks = st->nblocks; }
  if (first_block < 0 || first_block + nblocks > st->nblocks) return fail(CLOUDSC_B200_EARG, "step: block range out of bounds");
  FieldTable ft;
  for (int i = 0; i < CLOUDSC_NFIELDS; ++i) ft.p[i] = st->dev[i];
  long long end = (long long)(first_block + nblocks) * st->nproma;
  if (end > st->ngptot) end = st->ngptot;
  LaunchGeom g{first_block * st->nproma, (int)end, st->nproma, st->klev};
  if (st->has_tm && g_variant == VARIANT_AUTO)
    CU(launch_sweep_tmap(ft, st->tm, g, ptsphy, st->precision, (cudaStream_t)stream));
  else
    CU(launch_sweep(ft, g, ptsphy, st->precision, g_variant, (cudaStream_t)stream));
  return 0;
}

int cloudsc_b200_state_kernel_info(cloudsc_b200_state *st, int out[8]) {
  if (!st || !out) return fail(CLOUDSC_B200_EARG, "state_kernel_info: NULL argument");
  if (st->has_tm && g_variant == VARIANT_AUTO) CU(tmap_kernel_info(st->precision, out));
  else CU(sweep_kernel_info(st->nproma, st->precision, g_variant, out));
  return 0;
}

void *cloudsc_b200_field_ptr(cloudsc_b200_state *st, int f) {
  return (st && f >= 0 && f < CLOUDSC_NFIELDS) ? st->dev[f] : nullptr;
}

/* ---------------------------------------------------------------------------------------------- */
/* host-resident fields: chunked multi-stream pipeline                                            */
/* ---------------------------------------------------------------------------------------------- */
namespace {
/*
 * Per-device pipeline resources, kept between calls: cudaMalloc/cudaFree cost ~1 ms / ~0.4 ms each on
 * B200 (tools/ubench/pcie.cu), which for 51 fields x 3 lanes would be twice the time the copies
 * themselves need.  One arena per lane, carved into the live fields at 256-byte boundaries.
 */
constexpr int kMaxLanes = 8;
struct Lane {
  cudaStream_t stream = nullptr;
  cudaEvent_t k0 = nullptr, k1 = nullptr;
  char *arena = nullptr;
  bool used = false;
};
struct Pipe {
  std::mutex mu;
  int nlanes = 0;
  size_t lane_bytes = 0;
  Lane lanes[kMaxLanes];
  cudaEvent_t w0 = nullptr, w1 = nullptr;
  void release() {
    for (int i = 0; i < kMaxLanes; ++i) {
      Lane &l = lanes[i];
      if (l.arena) cudaFree(l.arena);
      if (l.k0) cudaEventDestroy(l.k0);
      if (l.k1) cudaEventDestroy(l.k1);
      if (l.stream) cudaStreamDestroy(l.stream);
      l = Lane();
    }
    if (w0) cudaEventDestroy(w0);
    if (w1) cudaEventDestroy(w1);
    w0 = w1 = nullptr;
    nlanes = 0;
    lane_bytes = 0;
  }
};
Pipe g_pipes[64];

cudaError_t pipe_reserve(Pipe &p, int nlanes, size_t lane_bytes) {
  cudaError_t e;
  if (!p.w0) {
    if ((e = cudaEventCreate(&p.w0)) != cudaSuccess) return e;
    if ((e = cudaEventCreate(&p.w1)) != cudaSuccess) return e;
  }
  for (int i = 0; i < nlanes; ++i) {
    Lane &l = p.lanes[i];
    if (!l.stream) {
      if ((e = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
      if ((e = cudaEventCreate(&l.k0)) != cudaSuccess) return e;
      if ((e = cudaEventCreate(&l.k1)) != cudaSuccess) return e;
    }
    if (!l.arena || lane_bytes > p.lane_bytes) {
      if (l.arena) cudaFree(l.arena);
      l.arena = nullptr;
      if ((e = cudaMalloc(&l.arena, lane_bytes)) != cudaSuccess) return e;
    }
    l.used = false;
  }
  if (lane_bytes > p.lane_bytes) {
    /* lanes beyond nlanes keep smaller arenas: drop them so that lane_bytes describes every arena */
    for (int i = nlanes; i < kMaxLanes; ++i)
      if (p.lanes[i].arena) { cudaFree(p.lanes[i].arena); p.lanes[i].arena = nullptr; }
    p.lane_bytes = lane_bytes;
  } else {
    for (int i = 0; i < nlanes; ++i)
      if (!p.lanes[i].arena) return cudaErrorMemoryAllocation;
  }
  if (nlanes > p.nlanes) p.nlanes = nlanes;
  return cudaSuccess;
}
}  // namespace

int cloudsc_b200_release_cache(void) {
  for (auto &p : g_pipes) {
    std::lock_guard<std::mutex> lk(p.mu);
    p.release();
  }
  return 0;
}

int cloudsc_b200_run_host(void *const host[CLOUDSC_NFIELDS], int ngptot, int nproma, int klev, double ptsphy,
                          int chunk_blocks, int nstreams, double stats[4]) {
  if (int rc = check_geom(ngptot, nproma, klev)) return rc;
  if (int rc = check_fields(host)) return rc;
  if (int rc = current_device_ready()) return rc;
  const int nblocks = (ngptot + nproma - 1) / nproma;
  if (nstreams <= 0) nstreams = 4;
  if (nstreams > kMaxLanes) nstreams = kMaxLanes;
  if (chunk_blocks <= 0) {
    /* ~6 chunks per lane: pipeline fill (first upload) and drain (last kernel + download) stay
       below ~5 % of the transfer time.  The kernel is ~2.5 % of a chunk's copy time, so chunks are
       sized for overlap, not for kernel occupancy; the floor only avoids degenerate launches. */
    chunk_blocks = (nblocks + nstreams * 6 - 1) / (nstreams * 6);
    const int min_blocks = (4096 + nproma - 1) / nproma;
    if (chunk_blocks < min_blocks) chunk_blocks = min_blocks;
  }
  if (chunk_blocks > nblocks) chunk_blocks = nblocks;
  const int nchunks = (nblocks + chunk_blocks - 1) / chunk_blocks;
  if (nstreams > nchunks) nstreams = nchunks;

  /* arena layout of one lane */
  size_t off[CLOUDSC_NFIELDS], blk[CLOUDSC_NFIELDS], lane_bytes = 0;
  for (int f = 0; f < CLOUDSC_NFIELDS; ++f) {
    off[f] = 0;
    blk[f] = rows_of(f, klev) * nproma * elem_size(f, CLOUDSC_B200_FP64);
    if (!is_live(f)) continue;
    off[f] = lane_bytes;
    lane_bytes += ((size_t)chunk_blocks * blk[f] + 255) / 256 * 256;
  }

  int dev = 0;
  CU(cudaGetDevice(&dev));
  Pipe &pipe = g_pipes[dev & 63];
  std::lock_guard<std::mutex> lk(pipe.mu);
  {
    cudaError_t e = pipe_reserve(pipe, nstreams, lane_bytes);
    if (e != cudaSuccess) { pipe.release(); return cuda_fail(e, "run_host: pipeline resources"); }
  }

  double kernel_ms = 0.0;
  size_t h2d = 0, d2h = 0;
  /* on error: drain what is in flight, then report (the cached resources stay valid) */
#define CUX(call)                                                    \
  do {                                                               \
    cudaError_t e_ = (call);                                         \
    if (e_ != cudaSuccess) {                                         \
      for (int i_ = 0; i_ < nstreams; ++i_) cudaStreamSynchronize(pipe.lanes[i_].stream); \
      return cuda_fail(e_, #call);                                   \
    }                                                                \
  } while (0)

  CUX(cudaEventRecord(pipe.w0, 0));
  for (int c = 0; c < nchunks; ++c) {
    Lane &l = pipe.lanes[c % nstreams];
    const int b0 = c * chunk_blocks;
    const int nb = (b0 + chunk_blocks <= nblocks) ? chunk_blocks : nblocks - b0;
    if (l.used) {  /* harvest the kernel time of the chunk that used this lane before */
      CUX(cudaEventSynchronize(l.k1));
      float ms = 0.f;
      CUX(cudaEventElapsedTime(&ms, l.k0, l.k1));
      kernel_ms += ms;
    }
    FieldTable ft;
    for (int f = 0; f < CLOUDSC_NFIELDS; ++f) ft.p[f] = is_live(f) ? l.arena + off[f] : nullptr;
    for (int f = 0; f < CLOUDSC_NFIELDS; ++f) {
      if (!is_in(f)) continue;
      CUX(cudaMemcpyAsync(ft.p[f], (const char *)host[f] + (size_t)b0 * blk[f], (size_t)nb * blk[f], cudaMemcpyHostToDevice, l.stream));
      h2d += (size_t)nb * blk[f];
    }
    long long cend = (long long)(b0 + nb) * nproma;
    if (cend > ngptot) cend = ngptot;
    LaunchGeom g{0, (int)(cend - (long long)b0 * nproma), nproma, klev};
    CUX(cudaEventRecord(l.k0, l.stream));
    CUX(launch_sweep(ft, g, ptsphy, CLOUDSC_B200_FP64, g_variant, l.stream));
    CUX(cudaEventRecord(l.k1, l.stream));
    for (int f = 0; f < CLOUDSC_NFIELDS; ++f) {
      if (!is_out(f)) continue;
      CUX(cudaMemcpyAsync((char *)host[f] + (size_t)b0 * blk[f], ft.p[f], (size_t)nb * blk[f], cudaMemcpyDeviceToHost, l.stream));
      d2h += (size_t)nb * blk[f];
    }
    l.used = true;
  }
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
#undef CUX
  return 0;
}

/* ---------------------------------------------------------------------------------------------- */
/* drop-in cloudsc_c(): one NPROMA block, host pointers, reference globals                        */
/* ---------------------------------------------------------------------------------------------- */
namespace {
struct ShimBuf {
  void *dev[CLOUDSC_NFIELDS] = {nullptr};
  int klon = 0, klev = 0, device = -1;
  cudaStream_t stream = nullptr;
  const struct TECLDP *seen_ptr = nullptr;
  struct TECLDP seen_val;
  struct cloudsc_consts seen_k;
};
thread_local ShimBuf t_shim;
}  // namespace

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
  if (int rc = check_fields(host)) return rc;
  if (!yrecldp) return fail(CLOUDSC_B200_EARG, "cloudsc_c: yrecldp is NULL");

  ShimBuf &s = t_shim;
  int dev = 0;
  CU(cudaGetDevice(&dev));
  /* re-snapshot the globals whenever they changed since this thread's last call */
  struct cloudsc_consts k;
#define CLOUDSC_X(n) k.n = n;
  CLOUDSC_KERNEL_CONSTS(CLOUDSC_X)
#undef CLOUDSC_X
  bool ready;
  { std::lock_guard<std::mutex> lk(g_mu); ready = dev >= 0 && dev < 64 && g_params_set[dev]; }
  if (!ready || s.device != dev || memcmp(&s.seen_val, yrecldp, sizeof(struct TECLDP)) != 0 || memcmp(&s.seen_k, &k, sizeof(k)) != 0) {
    if (int rc = cloudsc_b200_set_params(yrecldp, &k)) return rc;
    s.seen_val = *yrecldp;
    s.seen_k = k;
  }
  if (s.device != dev || s.klon != klon || s.klev != klev) {
    for (int f = 0; f < CLOUDSC_NFIELDS; ++f)
      if (s.dev[f]) { cudaFree(s.dev[f]); s.dev[f] = nullptr; }
    for (int f = 0; f < CLOUDSC_NFIELDS; ++f) {
      if (!is_live(f)) continue;
      CU(cudaMalloc(&s.dev[f], rows_of(f, klev) * klon * elem_size(f, CLOUDSC_B200_FP64)));
    }
    if (!s.stream) CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    s.device = dev; s.klon = klon; s.klev = klev;
  }
  const int ncols = kfdia - kidia + 1;
  for (int f = 0; f < CLOUDSC_NFIELDS; ++f) {
    if (!is_in(f)) continue;
    const size_t es = elem_size(f, CLOUDSC_B200_FP64);
    CU(cudaMemcpy2DAsync((char *)s.dev[f] + (size_t)(kidia - 1) * es, (size_t)klon * es,
                         (const char *)host[f] + (size_t)(kidia - 1) * es, (size_t)klon * es,
                         (size_t)ncols * es, rows_of(f, klev), cudaMemcpyHostToDevice, s.stream));
  }
  FieldTable ft;
  for (int i = 0; i < CLOUDSC_NFIELDS; ++i) ft.p[i] = s.dev[i];
  LaunchGeom g{kidia - 1, kfdia, klon, klev};
  CU(launch_sweep(ft, g, ptsphy, CLOUDSC_B200_FP64, g_variant, s.stream));
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
