/*
 * cloudsc_kernels.cu -- sm_100a kernels of the CLOUDSC hot path.
 *
 * Replaces the OpenMP block loop + per-block call of the reference driver
 * (src/cloudsc_c/cloudsc/cloudsc_driver.c:183-229 -> cloudsc_c.c:19) by ONE launch: one thread per
 * column, the whole 137-level sweep in registers (cloudsc_column.cuh), parameters in __constant__
 * memory.  Field layout is the driver's [block][level][column] with NPROMA fastest, so a warp's 32
 * columns of one (field, level) are 256 contiguous bytes: every global access is fully coalesced.
 */
#include "cloudsc_kernels.cuh"
#include "cloudsc_column.cuh"
#include "cloudsc_fill.h"

namespace cloudsc {

__constant__ CloudscConst<double> g_const64;
__constant__ CloudscConst<float> g_const32;

template <typename T> __device__ __forceinline__ const CloudscConst<T> &dev_consts();
template <> __device__ __forceinline__ const CloudscConst<double> &dev_consts<double>() { return g_const64; }
template <> __device__ __forceinline__ const CloudscConst<float> &dev_consts<float>() { return g_const32; }

/* ---- slot -> field-table mapping (compile-time) ----------------------------------------------- */
__host__ __device__ constexpr int in_field(int s) {
  return s == IN_PT ? CLOUDSC_F_PT : s == IN_PQ ? CLOUDSC_F_PQ : s == IN_TT ? CLOUDSC_F_TENDENCY_TMP_T
       : s == IN_TQ ? CLOUDSC_F_TENDENCY_TMP_Q : s == IN_TA ? CLOUDSC_F_TENDENCY_TMP_A : s == IN_PA ? CLOUDSC_F_PA
       : s == IN_PAP ? CLOUDSC_F_PAP : s == IN_PSUPSAT ? CLOUDSC_F_PSUPSAT : s == IN_PLUDE ? CLOUDSC_F_PLUDE
       : s == IN_PLU ? CLOUDSC_F_PLU : s == IN_PSNDE ? CLOUDSC_F_PSNDE : s == IN_PMFU ? CLOUDSC_F_PMFU
       : s == IN_PMFD ? CLOUDSC_F_PMFD : s == IN_PHRSW ? CLOUDSC_F_PHRSW : s == IN_PHRLW ? CLOUDSC_F_PHRLW
       : s == IN_PVERVEL ? CLOUDSC_F_PVERVEL : s == IN_PVFL ? CLOUDSC_F_PVFL : s == IN_PVFI ? CLOUDSC_F_PVFI
       : (s >= IN_PCLV0 && s <= IN_PCLV3) ? CLOUDSC_F_PCLV
       : (s >= IN_TCLD0 && s <= IN_TCLD3) ? CLOUDSC_F_TENDENCY_TMP_CLD : CLOUDSC_F_PAPH;
}
__host__ __device__ constexpr int in_species(int s) {
  return (s >= IN_PCLV0 && s <= IN_PCLV3) ? s - IN_PCLV0 : (s >= IN_TCLD0 && s <= IN_TCLD3) ? s - IN_TCLD0 : 0;
}
/* 0: [klev] field, 1: [klev+1] field, 2: [nclv][klev] field */
__host__ __device__ constexpr int in_kind(int s) { return s == IN_PAPH ? 1 : (s >= IN_PCLV0 && s <= IN_TCLD3) ? 2 : 0; }

__host__ __device__ constexpr int out_field(int s) {
  return s == OUT_TT ? CLOUDSC_F_TENDENCY_LOC_T : s == OUT_TQ ? CLOUDSC_F_TENDENCY_LOC_Q
       : s == OUT_TA ? CLOUDSC_F_TENDENCY_LOC_A : (s >= OUT_TCLD0 && s <= OUT_TCLD4) ? CLOUDSC_F_TENDENCY_LOC_CLD
       : s == OUT_COVPTOT ? CLOUDSC_F_PCOVPTOT : s == OUT_PLUDE ? CLOUDSC_F_PLUDE
       : s == OUT_FSQLF ? CLOUDSC_F_PFSQLF : s == OUT_FSQIF ? CLOUDSC_F_PFSQIF : s == OUT_FCQNNG ? CLOUDSC_F_PFCQNNG
       : s == OUT_FCQLNG ? CLOUDSC_F_PFCQLNG : s == OUT_FSQRF ? CLOUDSC_F_PFSQRF : s == OUT_FSQSF ? CLOUDSC_F_PFSQSF
       : s == OUT_FCQRNG ? CLOUDSC_F_PFCQRNG : s == OUT_FCQSNG ? CLOUDSC_F_PFCQSNG
       : s == OUT_FSQLTUR ? CLOUDSC_F_PFSQLTUR : s == OUT_FSQITUR ? CLOUDSC_F_PFSQITUR
       : s == OUT_FPLSL ? CLOUDSC_F_PFPLSL : s == OUT_FPLSN ? CLOUDSC_F_PFPLSN
       : s == OUT_FHPSL ? CLOUDSC_F_PFHPSL : CLOUDSC_F_PFHPSN;
}
__host__ __device__ constexpr int out_species(int s) { return (s >= OUT_TCLD0 && s <= OUT_TCLD4) ? s - OUT_TCLD0 : 0; }
__host__ __device__ constexpr int out_kind(int s) { return s >= OUT_FSQLF ? 1 : (s >= OUT_TCLD0 && s <= OUT_TCLD4) ? 2 : 0; }

/* ---- IO policy 1: straight global memory ------------------------------------------------------ */
/*
 * Addresses are rebuilt per access from the kernel-parameter table (constant bank) plus three
 * per-thread element offsets, instead of keeping ~50 offset pointers alive in registers.
 * Inputs go through the read-only path (ld.global.nc); PLUDE is in/out and uses a plain load.
 * Outputs are streaming stores (st.global.cs): nothing written is ever re-read by this kernel.
 */
template <typename T>
struct GlobalIO {
  const FieldTable &ft;
  long long offL, offH, offC; /* element offset of (block, level 0, column) for the 3 field kinds */
  long long off1;             /* block*nproma + jl */
  int stride;                 /* nproma */
  int klev;

  template <int S> __device__ __forceinline__ T ld(int k) const {
    const T *base = static_cast<const T *>(ft.p[in_field(S)]);
    const long long off = (in_kind(S) == 0 ? offL : in_kind(S) == 1 ? offH : offC + (long long)in_species(S) * klev * stride)
                        + (long long)k * stride;
    if (S == IN_PLUDE) return base[off];
    return __ldg(base + off);
  }
  template <int S> __device__ __forceinline__ void st(int k, T v) {
    T *base = static_cast<T *>(ft.p[out_field(S)]);
    const long long off = (out_kind(S) == 0 ? offL : out_kind(S) == 1 ? offH : offC + (long long)out_species(S) * klev * stride)
                        + (long long)k * stride;
    __stcs(base + off, v);
  }
  __device__ __forceinline__ T plsm() const { return __ldg(static_cast<const T *>(ft.p[CLOUDSC_F_PLSM]) + off1); }
  __device__ __forceinline__ int ktype() const { return __ldg(static_cast<const int *>(ft.p[CLOUDSC_F_KTYPE]) + off1); }
  __device__ __forceinline__ void st_rainfrac(T v) { static_cast<T *>(ft.p[CLOUDSC_F_PRAINFRAC_TOPRFZ])[off1] = v; }
  __device__ __forceinline__ void level_begin(int) {}
  __device__ __forceinline__ void level_end(int) {}
};

template <typename T, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
cloudsc_sweep_direct(const __grid_constant__ FieldTable ft, const LaunchGeom g, const __grid_constant__ StepConst<T> sc) {
  const int col = g.col_begin + blockIdx.x * BLOCK + threadIdx.x;
  if (col >= g.col_end) return;
  const int ibl = col / g.nproma;
  const int jl = col - ibl * g.nproma;
  GlobalIO<T> io{ft,
                 (long long)ibl * g.klev * g.nproma + jl,
                 (long long)ibl * (g.klev + 1) * g.nproma + jl,
                 (long long)ibl * 5 * g.klev * g.nproma + jl,
                 (long long)ibl * g.nproma + jl,
                 g.nproma, g.klev};
  cloudsc_column<T>(dev_consts<T>(), sc, io, g.klev);
}

/* ---- host side -------------------------------------------------------------------------------- */
/* host mirrors of the constant blocks of the current device (per-launch constants derive from them) */
static CloudscConst<double> h_const64[64];
static CloudscConst<float> h_const32[64];
template <typename T> static const CloudscConst<T> &host_consts(int dev);
template <> const CloudscConst<double> &host_consts<double>(int dev) { return h_const64[dev & 63]; }
template <> const CloudscConst<float> &host_consts<float>(int dev) { return h_const32[dev & 63]; }

cudaError_t upload_consts(const CloudscConst<double> &c64, const CloudscConst<float> &c32) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbol(g_const64, &c64, sizeof(c64));
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbol(g_const32, &c32, sizeof(c32));
  if (e != cudaSuccess) return e;
  h_const64[dev & 63] = c64;
  h_const32[dev & 63] = c32;
  return cudaSuccess;
}

namespace {
constexpr int kDirectBlock = 128;

/* register budget of the direct kernel = 65536 / (128 * MINB): 2 -> 255, 3 -> 168, 4 -> 128 */
inline int direct_minb(int variant) {
  switch (variant) {
    case VARIANT_DIRECT_R255: return 2;
    case VARIANT_DIRECT_R128: return 4;
    case VARIANT_DIRECT_R168:
    default: return 3;
  }
}

template <typename T, int MINB>
cudaError_t launch_direct_mb(const FieldTable &ft, const LaunchGeom &g, double ptsphy, cudaStream_t stream) {
  const int ncols = g.col_end - g.col_begin;
  const int grid = (ncols + kDirectBlock - 1) / kDirectBlock;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  StepConst<T> sc;
  fill_step<T>(host_consts<T>(dev), ptsphy, &sc);
  cloudsc_sweep_direct<T, kDirectBlock, MINB><<<grid, kDirectBlock, 0, stream>>>(ft, g, sc);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_direct(const FieldTable &ft, const LaunchGeom &g, double ptsphy, int variant, cudaStream_t stream) {
  switch (direct_minb(variant)) {
    case 2: return launch_direct_mb<T, 2>(ft, g, ptsphy, stream);
    case 4: return launch_direct_mb<T, 4>(ft, g, ptsphy, stream);
    default: return launch_direct_mb<T, 3>(ft, g, ptsphy, stream);
  }
}

template <typename T>
const void *direct_fn(int variant) {
  switch (direct_minb(variant)) {
    case 2: return (const void *)cloudsc_sweep_direct<T, kDirectBlock, 2>;
    case 4: return (const void *)cloudsc_sweep_direct<T, kDirectBlock, 4>;
    default: return (const void *)cloudsc_sweep_direct<T, kDirectBlock, 3>;
  }
}
}  // namespace

cudaError_t launch_sweep(const FieldTable &ft, const LaunchGeom &g, double ptsphy, int precision, int variant,
                         cudaStream_t stream) {
  if (g.col_end <= g.col_begin) return cudaSuccess;
  if (precision == CLOUDSC_B200_FP32) return launch_direct<float>(ft, g, ptsphy, variant, stream);
  return launch_direct<double>(ft, g, ptsphy, variant, stream);
}

cudaError_t sweep_kernel_info(int nproma, int precision, int variant, int out[8]) {
  (void)nproma;
  cudaFuncAttributes a;
  const void *fn = precision == CLOUDSC_B200_FP32 ? direct_fn<float>(variant) : direct_fn<double>(variant);
  cudaError_t e = cudaFuncGetAttributes(&a, fn);
  if (e != cudaSuccess) return e;
  int nb = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, kDirectBlock, 0);
  if (e != cudaSuccess) return e;
  out[0] = a.numRegs;
  out[1] = (int)a.localSizeBytes;
  out[2] = (int)a.sharedSizeBytes;
  out[3] = a.maxThreadsPerBlock;
  out[4] = nb;
  out[5] = kDirectBlock;
  out[6] = variant;
  out[7] = 0;
  return cudaSuccess;
}

}  // namespace cloudsc
