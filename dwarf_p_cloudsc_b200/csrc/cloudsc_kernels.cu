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
#include "cloudsc_split_plan.h"

#include <atomic>
#include <cstdint>
#include <cstdlib>

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
  template <int S> __device__ __forceinline__ T ld_next(int k) const { return ld<S>(k + 1); }
  /* aerosol inputs ([klev] fields), only read by the AER instantiation where the matching flag is set */
  template <int A> __device__ __forceinline__ T ld_aer(int k) const {
    const T *base = static_cast<const T *>(ft.p[A == AER_RE_ICE ? CLOUDSC_F_PRE_ICE : A == AER_ICRIT ? CLOUDSC_F_PICRIT_AER : CLOUDSC_F_PNICE]);
    return __ldg(base + offL + (long long)k * stride);
  }
  __device__ __forceinline__ T paph_row(int r) const { return ld<IN_PAPH>(r); }
  template <int S> __device__ __forceinline__ void st(int k, T v) {
    T *base = static_cast<T *>(ft.p[out_field(S)]);
    const long long off = (out_kind(S) == 0 ? offL : out_kind(S) == 1 ? offH : offC + (long long)out_species(S) * klev * stride)
                        + (long long)k * stride;
    __stcs(base + off, v);
  }
  __device__ __forceinline__ T plsm() const { return __ldg(static_cast<const T *>(ft.p[CLOUDSC_F_PLSM]) + off1); }
  __device__ __forceinline__ int ktype() const { return __ldg(static_cast<const int *>(ft.p[CLOUDSC_F_KTYPE]) + off1); }
  __device__ __forceinline__ void st_rainfrac(T v) { static_cast<T *>(ft.p[CLOUDSC_F_PRAINFRAC_TOPRFZ])[off1] = v; }
  template <int S> __device__ __forceinline__ void st_row0(T v) { st<S>(0, v); }
  __device__ __forceinline__ void level_begin(int) {}
  __device__ __forceinline__ void level_end(int) {}
};

template <typename T, int BLOCK, int MINB, bool AER>
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
  cloudsc_column<T, AER>(dev_consts<T>(), sc, io, g.klev);
}

/* ---- IO policy 2: TMA-staged input tiles shared by the CTA, streaming stores ------------------- */
/*
 * A CTA of WARPS warps owns 32*WARPS consecutive columns of ONE NPROMA block, so one (field, level)
 * row of the CTA is 32*WARPS*sizeof(T) contiguous, 16-byte aligned bytes: exactly one
 * `cp.async.bulk` (TMA; SASS UBLKCP).
 *
 * INPUTS.  Per level the column code needs N_IN = 27 rows.  They are fetched TWO LEVELS AHEAD into
 * a 2-stage shared-memory ring, signalled through one "full" mbarrier per stage:
 *   level_begin(k): every thread waits on the phase of full[k&1];
 *   ld<S>(k):       ONE LDS at [stage + S*row + tid] -- no global address arithmetic in the consumer;
 *   level_end(k):   lane 0 of every warp arrives on empty[k&1]; thread 0 waits until all WARPS warps
 *                   have arrived (the stage is no longer read), arms full[k&1] with expect_tx and
 *                   issues the 27 bulk copies of level k+2.  All of thread 0's operands derive from
 *                   blockIdx / kernel parameters, so the copy loop runs on the uniform datapath.
 * Rows of the fields read at level k+1 while level k is processed (paph, pmfu, pmfd, plu) are stored
 * in the stage of level k, so a stage holds exactly what one iteration consumes.  A whole level of
 * compute (~10 us per warp) lies between issue and use: the long-scoreboard stalls of the direct
 * kernel disappear without spending registers on prefetch.
 *
 * OUTPUTS go straight from registers to global memory as streaming stores (st.global.cs); stores do
 * not stall the issuing warp, and keeping them out of shared memory leaves room for 12-16 resident
 * warps per SM (2 stages x 27 rows x 1 KB = 54 KB per 4-warp CTA).
 */
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

/* one lane of a converged warp (elect.sync): lets ptxas issue the guarded uniform-datapath code once
 * instead of wrapping every bulk copy in a "for each distinct active lane" loop */
__device__ __forceinline__ bool elect_one() {
  unsigned pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

/* 1 for the input slots whose row k+1 is staged with level k */
__host__ __device__ constexpr int in_next(int s) { return (s == IN_PAPH || s == IN_PMFU || s == IN_PMFD || s == IN_PLU) ? 1 : 0; }

template <typename T, int WARPS>
struct TileIO {
  static constexpr int kCols = 32 * WARPS;                 /* columns of the CTA tile */
  static constexpr int kRow = kCols * (int)sizeof(T);      /* bytes of one staged row */
  static constexpr int kStage = N_IN * kRow;               /* one input stage */
  static constexpr int kScratch = PARK_COUNT;              /* rows of per-thread scratch behind the ring (park_ld/park_st) */
  static constexpr int kSmem = 2 * kStage + 32 + kScratch * kRow;   /* ring + full[2], empty[2] mbarriers + scratch */
  const FieldTable &ft;
  /* CTA-uniform: element offsets of (block, level 0, first column of the tile) per field kind */
  long long tileL, tileH, tileC;
  int stride, klev;                   /* nproma, klev */
  unsigned smem0;                     /* shared address of stage 0 */
  unsigned bar0;                      /* shared address of full[0]; full[1] +8, empty[0] +16, empty[1] +24 */
  /* per thread */
  const T *stage_t;                   /* generic pointer: stage 0 + tid */
  const T *cur;                       /* stage of the level being processed + tid */
  long long offH, off1;               /* element offsets of this thread's column: paph rows, per-column scalars */
  /* 32-bit element offsets of the CURRENT level's output rows (launcher guarantees every field has
     fewer than 2^32 elements): base + offset is one IMAD.WIDE.U32 per store.  cH1 is row k+1 of the
     [klev+1] flux fields, which is what level k writes. */
  unsigned cL, cH1, cC, spstride;
  /* false for the lanes of the last tile that lie beyond col_end (ragged NGPTOT): they run the sweep on the padding
     columns of the last NPROMA block -- the same kernel, hence the same rounding, for every column -- but never store */
  bool live;

  __device__ __forceinline__ void issue(int level) {   /* thread 0 only */
    const unsigned st = (unsigned)level & 1u;
    const unsigned bar = bar0 + 8u * st;
    const unsigned dst = smem0 + st * (unsigned)kStage;
    mbar_arrive_expect_tx(bar, (unsigned)kStage);
    const long long r0 = (long long)level * stride;
    /* pmfu/pmfd/plu have klev rows: at the last level their "next" row does not exist and is not
       read; re-fetch the last row to keep the transaction count of the stage constant */
    const long long r1 = (long long)(level + 1 < klev ? level + 1 : klev - 1) * stride;
    const long long r1h = (long long)(level + 1) * stride;
#pragma unroll
    for (int s = 0; s < N_IN; ++s) {
      const long long tile = in_kind(s) == 0 ? tileL : in_kind(s) == 1 ? tileH : tileC + (long long)in_species(s) * klev * stride;
      const long long row = s == IN_PAPH ? r1h : in_next(s) ? r1 : r0;
      bulk_g2s(dst + (unsigned)(s * kRow), static_cast<const T *>(ft.p[in_field(s)]) + (tile + row), (unsigned)kRow, bar);
    }
  }
  __device__ __forceinline__ void level_begin(int k) {
    cur = stage_t + (k & 1) * (kStage / (int)sizeof(T));
    mbar_wait(bar0 + 8u * ((unsigned)k & 1u), ((unsigned)k >> 1) & 1u);
  }
  __device__ __forceinline__ void level_end(int k) {
    __syncwarp();               /* every lane of this warp is done reading stage k&1 */
    if (WARPS > 1 && (threadIdx.x & 31) == 0) mbar_arrive(bar0 + 16u + 8u * ((unsigned)k & 1u));
    if (k + 2 < klev && (WARPS == 1 || threadIdx.x < 32)) {   /* warp-uniform: the whole warp 0 enters */
      if (elect_one()) {
        if (WARPS > 1) mbar_wait(bar0 + 16u + 8u * ((unsigned)k & 1u), ((unsigned)k >> 1) & 1u);
        issue(k + 2);
      }
      __syncwarp();
    }
    cL += (unsigned)stride; cH1 += (unsigned)stride; cC += (unsigned)stride;
  }
  __device__ __forceinline__ T &scr(int i) const {
    return const_cast<T *>(stage_t)[(2 * kStage + 32) / (int)sizeof(T) + i * kCols];
  }
  template <int S> __device__ __forceinline__ T ld(int) const { return cur[S * kCols]; }
  template <int S> __device__ __forceinline__ T ld_next(int) const { return cur[S * kCols]; }
  __device__ __forceinline__ T paph_row(int r) const {
    return __ldg(static_cast<const T *>(ft.p[CLOUDSC_F_PAPH]) + offH + (long long)r * stride);
  }
  template <int S> __device__ __forceinline__ void st(int, T v) {
    const unsigned off = out_kind(S) == 0 ? cL : out_kind(S) == 1 ? cH1 : cC + (unsigned)out_species(S) * spstride;
    if (live) __stcs(static_cast<T *>(ft.p[out_field(S)]) + off, v);
  }
  /* row 0 of a flux field, before the level loop (cH1 still points at row 1) */
  template <int S> __device__ __forceinline__ void st_row0(T v) {
    if (live) __stcs(static_cast<T *>(ft.p[out_field(S)]) + (cH1 - (unsigned)stride), v);
  }
  __device__ __forceinline__ T plsm() const { return __ldg(static_cast<const T *>(ft.p[CLOUDSC_F_PLSM]) + off1); }
  __device__ __forceinline__ int ktype() const { return __ldg(static_cast<const int *>(ft.p[CLOUDSC_F_KTYPE]) + off1); }
  __device__ __forceinline__ void st_rainfrac(T v) { if (live) static_cast<T *>(ft.p[CLOUDSC_F_PRAINFRAC_TOPRFZ])[off1] = v; }
};

template <typename T, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
cloudsc_sweep_tile(const __grid_constant__ FieldTable ft, const LaunchGeom g, const __grid_constant__ StepConst<T> sc) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using IO = TileIO<T, WARPS>;
  /* first column of this CTA; col_begin and nproma are multiples of 32*WARPS (launcher); col_end may be ragged */
  const int ccol = g.col_begin + blockIdx.x * IO::kCols;
  if (ccol >= g.col_end) return;  /* CTA-uniform */
  const int ibl = ccol / g.nproma;
  const int jl0 = ccol - ibl * g.nproma;
  const int jl = jl0 + (int)threadIdx.x;
  const long long bL = (long long)ibl * g.klev * g.nproma, bH = (long long)ibl * (g.klev + 1) * g.nproma;

  IO io{ft, bL + jl0, bH + jl0, bL * 5 + jl0, g.nproma, g.klev,
        smem_u32(smem_raw), smem_u32(smem_raw + 2 * IO::kStage),
        reinterpret_cast<const T *>(smem_raw) + threadIdx.x,
        nullptr, bH + jl, (long long)ibl * g.nproma + jl,
        (unsigned)(bL + jl), (unsigned)(bH + jl + g.nproma), (unsigned)(bL * 5 + jl), (unsigned)(g.klev * g.nproma),
        ccol + (int)threadIdx.x < g.col_end};
  if (threadIdx.x == 0) {
    mbar_init(io.bar0, 1);
    mbar_init(io.bar0 + 8, 1);
    mbar_init(io.bar0 + 16, WARPS);
    mbar_init(io.bar0 + 24, WARPS);
    mbar_fence_init();
  }
  __syncthreads();
  if (WARPS == 1 || threadIdx.x < 32) {
    if (elect_one()) {
      io.issue(0);
      if (g.klev > 1) io.issue(1);
    }
    __syncwarp();
  }
  cloudsc_column<T>(dev_consts<T>(), sc, io, g.klev);
}

/* ---- IO policy 4: tensor-map TMA on an arena-allocated input set --------------------------------- */
/*
 * Issuing a bulk copy costs the issuing warp ~50 cycles whatever its size (measured: 27 copies per level
 * = 1500-5000 cycles of a 6000-15000-cycle level, tools/gpu_level_timing.py), so the number of requests
 * matters more than their bytes.  When the input fields come from ONE arena with a constant spacing
 * between fields of the same kind (cloudsc_b200_create), a 4-/5-dimensional tensor map describes
 * (column, level, [species,] block, field) and one `cp.async.bulk.tensor` request brings the rows of
 * all fields of a group: 4 requests per level instead of 27.  Same 2-stage ring, same mbarriers, same
 * consumer (one LDS per value) as IO policy 2 with one-warp CTAs.
 */
/* group a: first the 7 fields every level reads, then the 8 that only the physics levels read (k >= ncldtop - 1; pap one
 * level earlier): above the cloud-top level one request with the 7-field box `a_top` replaces the 15-field one */
const int kArenaGroupA[15] = {CLOUDSC_F_PT, CLOUDSC_F_TENDENCY_TMP_T, CLOUDSC_F_TENDENCY_TMP_A, CLOUDSC_F_PA, CLOUDSC_F_PLUDE,
                              CLOUDSC_F_PVFL, CLOUDSC_F_PVFI,
                              CLOUDSC_F_PQ, CLOUDSC_F_TENDENCY_TMP_Q, CLOUDSC_F_PAP, CLOUDSC_F_PSUPSAT, CLOUDSC_F_PSNDE,
                              CLOUDSC_F_PHRSW, CLOUDSC_F_PHRLW, CLOUDSC_F_PVERVEL};
const int kArenaGroupB[3] = {CLOUDSC_F_PLU, CLOUDSC_F_PMFU, CLOUDSC_F_PMFD};
const int kArenaGroupC[2] = {CLOUDSC_F_PCLV, CLOUDSC_F_TENDENCY_TMP_CLD};
constexpr int kGroupATop = 7;   /* fields of group a read at every level */

/* shared-memory row of an input slot: group a rows 0-14 (kArenaGroupA order), b 15-17, c 18-25 ([field][species]), h 26 */
__host__ __device__ constexpr int tmap_row(int s) {
  return s == IN_PT ? 0 : s == IN_TT ? 1 : s == IN_TA ? 2 : s == IN_PA ? 3 : s == IN_PLUDE ? 4 : s == IN_PVFL ? 5 : s == IN_PVFI ? 6
       : s == IN_PQ ? 7 : s == IN_TQ ? 8 : s == IN_PAP ? 9 : s == IN_PSUPSAT ? 10 : s == IN_PSNDE ? 11 : s == IN_PHRSW ? 12
       : s == IN_PHRLW ? 13 : s == IN_PVERVEL ? 14
       : s == IN_PLU ? 15 : s == IN_PMFU ? 16 : s == IN_PMFD ? 17
       : (s >= IN_PCLV0 && s <= IN_PCLV3) ? 18 + (s - IN_PCLV0)
       : (s >= IN_TCLD0 && s <= IN_TCLD3) ? 22 + (s - IN_TCLD0) : 26;
}

__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *tm, unsigned bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap *tm, unsigned bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(unsigned dst, const CUtensorMap *tm, unsigned bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

/*
 * RANGED = false: one CTA sweeps levels [0, klev) of one tile (cloudsc_sweep_tmap).
 * RANGED = true:  the persistent kernel (cloudsc_sweep_persist): the same warp runs level ranges [kb, ke) of several
 *                 tiles one after the other; the ring's stage / mbarrier-phase bookkeeping continues across them
 *                 (`kofs` maps a level to the running count of levels this warp has staged), and the state at a cut
 *                 travels through `carry` (per tile: [kCarryVals][32 lanes] elements, coalesced).
 */
constexpr int kCarryRegs = 8;                        /* Carry members that live in registers in the parked code path */
constexpr int kCarryVals = kCarryRegs + PARK_COUNT;  /* + the parked slots */

template <typename T, bool RANGED = false>
struct TmapIO {
  static constexpr bool kRanged = RANGED;
  static constexpr int kCols = 32;
  static constexpr int kRow = kCols * (int)sizeof(T);
  static constexpr int kStage = N_IN * kRow;
  static constexpr int kScratch = PARK_COUNT;              /* rows of per-thread scratch behind the ring (park_ld/park_st) */
  static constexpr int kSmem = 2 * kStage + 32 + kScratch * kRow;
  const FieldTable &ft;
  const TensorMaps &tm;
  int ibl, jl0;                       /* NPROMA block and first column (inside it) of this warp */
  int stride, klev;
  int ktop1;                          /* ncldtop - 2: levels below it (k < ktop1) read the 7-field box only */
  unsigned smem0, bar0;               /* shared addresses of stage 0 and of full[0] (full[1] at +8) */
  const T *stage_t;                   /* generic pointer: stage 0 + lane */
  const T *cur;
  long long offH, off1;
  unsigned cL, cH1, cC, spstride;     /* 32-bit element offsets of the current level's output rows */
  bool live;                          /* see TileIO::live */
  /* RANGED only */
  int kb, ke, kofs;
  T *carry;                           /* this tile's [kCarryVals][32] block + lane */
  int n_ibl, n_jl0;                   /* streaming schedule: the tile this warp sweeps next (n_ibl < 0: none); its levels 0
                                         and 1 are requested during the last two levels of this tile, so that the ring never
                                         drains between tiles */

  __device__ __forceinline__ int it_of(int level) const { return RANGED ? level + kofs : level; }
  __device__ __forceinline__ int k_begin() const { return kb; }
  __device__ __forceinline__ int k_end() const { return ke; }

  __device__ __forceinline__ void issue(int level) { issue_tile(level, it_of(level), ibl, jl0); }   /* one elected lane */
  __device__ __forceinline__ void issue_tile(int level, int it, int ibl, int jl0) {
    const unsigned st = (unsigned)it & 1u;
    const unsigned bar = bar0 + 8u * st;
    const unsigned dst = smem0 + st * (unsigned)kStage;
    if (level < ktop1) {
      /* above the first physics level only 7 of the 15 group-a fields and none of group b are read (cloudsc_column.cuh:
         qv, the forcing terms and the convective fluxes enter at k >= ncldtop - 1, pap one level earlier); the rows not
         fetched keep stale values that no output depends on (tests: test_skipped_input_rows_are_dead) */
      mbar_arrive_expect_tx(bar, (unsigned)((kGroupATop + 8 + 1) * kRow));
      tma_load_4d(dst, &tm.a_top, bar, jl0, level, ibl, 0);
    } else {
      mbar_arrive_expect_tx(bar, (unsigned)kStage);
      tma_load_4d(dst, &tm.a, bar, jl0, level, ibl, 0);
      /* row level+1 of plu/pmfu/pmfd does not exist at the last level: the request is out of bounds there,
         TMA fills zeros (never consumed) and still reports the full box to the mbarrier */
      tma_load_4d(dst + 15u * kRow, &tm.b, bar, jl0, level + 1, ibl, 0);
    }
    tma_load_5d(dst + 18u * kRow, &tm.c, bar, jl0, level, 0, ibl, 0);
    tma_load_3d(dst + 26u * kRow, &tm.h, bar, jl0, level + 1, ibl);
  }
  /* -DCLOUDSC_TIMING (tools/build_alt.sh timing -DCLOUDSC_TIMING; tools/gpu_level_timing.py): four clock stamps
     per level of every 512th CTA into the TENDENCY_CML_T buffer, which the kernel never touches otherwise */
#ifdef CLOUDSC_TIMING
#define CLOUDSC_TSTAMP(i)                                                                        \
  if (!RANGED && threadIdx.x == 0 && blockIdx.x % 512 == 0 && ft.p[CLOUDSC_F_TENDENCY_CML_T])    \
    static_cast<long long *>(ft.p[CLOUDSC_F_TENDENCY_CML_T])[(blockIdx.x / 512) * 4 * klev + 4 * k + (i)] = clock64()
#else
#define CLOUDSC_TSTAMP(i)
#endif
  __device__ __forceinline__ void level_begin(int k) {
    const unsigned it = (unsigned)it_of(k);
    cur = stage_t + (it & 1u) * (kStage / (int)sizeof(T));
    CLOUDSC_TSTAMP(0);
    mbar_wait(bar0 + 8u * (it & 1u), (it >> 1) & 1u);
    CLOUDSC_TSTAMP(1);
  }
  __device__ __forceinline__ void level_end(int k) {
    CLOUDSC_TSTAMP(2);
    __syncwarp();               /* every lane is done reading stage k&1 */
    if (k + 2 < (RANGED ? ke : klev)) {
      if (elect_one()) issue(k + 2);
      __syncwarp();
    } else if (RANGED && n_ibl >= 0) {   /* (only with kb == 0, ke == klev: the streaming schedule) */
      if (elect_one()) issue_tile(k + 2 - ke, it_of(k + 2), n_ibl, n_jl0);
      __syncwarp();
    }
    cL += (unsigned)stride; cH1 += (unsigned)stride; cC += (unsigned)stride;
    CLOUDSC_TSTAMP(3);
  }
  __device__ __forceinline__ T &scr(int i) const {
    return const_cast<T *>(stage_t)[(2 * kStage + 32) / (int)sizeof(T) + i * kCols];
  }
  /* state at a cut (RANGED): the register members of Carry and every parked slot */
  __device__ __forceinline__ void save_carry(const Carry<T> &s) {
    T *c = carry;
    __stcg(c + 0 * 32, s.tp1_m1); __stcg(c + 1 * 32, s.a_m1); __stcg(c + 2 * 32, s.pap_m1); __stcg(c + 3 * 32, s.paph_k);
    __stcg(c + 4 * 32, s.mf_k); __stcg(c + 5 * 32, s.anewm1); __stcg(c + 6 * 32, s.qlnm1); __stcg(c + 7 * 32, s.qinm1);
#pragma unroll
    for (int i = 0; i < PARK_COUNT; ++i) __stcg(c + (kCarryRegs + i) * 32, scr(i));
  }
  __device__ __forceinline__ void load_carry(Carry<T> &s) {
    const T *c = carry;
    s.tp1_m1 = __ldcg(c + 0 * 32); s.a_m1 = __ldcg(c + 1 * 32); s.pap_m1 = __ldcg(c + 2 * 32); s.paph_k = __ldcg(c + 3 * 32);
    s.mf_k = __ldcg(c + 4 * 32); s.anewm1 = __ldcg(c + 5 * 32); s.qlnm1 = __ldcg(c + 6 * 32); s.qinm1 = __ldcg(c + 7 * 32);
#pragma unroll
    for (int i = 0; i < PARK_COUNT; ++i) scr(i) = __ldcg(c + (kCarryRegs + i) * 32);
  }
  template <int S> __device__ __forceinline__ T ld(int) const { return cur[tmap_row(S) * kCols]; }
  template <int S> __device__ __forceinline__ T ld_next(int) const { return cur[tmap_row(S) * kCols]; }
  __device__ __forceinline__ T paph_row(int r) const {
    return __ldg(static_cast<const T *>(ft.p[CLOUDSC_F_PAPH]) + offH + (long long)r * stride);
  }
  template <int S> __device__ __forceinline__ void st(int, T v) {
    const unsigned off = out_kind(S) == 0 ? cL : out_kind(S) == 1 ? cH1 : cC + (unsigned)out_species(S) * spstride;
    if (live) __stcs(static_cast<T *>(ft.p[out_field(S)]) + off, v);
  }
  template <int S> __device__ __forceinline__ void st_row0(T v) {
    if (live) __stcs(static_cast<T *>(ft.p[out_field(S)]) + (cH1 - (unsigned)stride), v);
  }
  __device__ __forceinline__ T plsm() const { return __ldg(static_cast<const T *>(ft.p[CLOUDSC_F_PLSM]) + off1); }
  __device__ __forceinline__ int ktype() const { return __ldg(static_cast<const int *>(ft.p[CLOUDSC_F_KTYPE]) + off1); }
  __device__ __forceinline__ void st_rainfrac(T v) { if (live) static_cast<T *>(ft.p[CLOUDSC_F_PRAINFRAC_TOPRFZ])[off1] = v; }
};

/* resident warps per SM the one-warp-CTA kernels are sized for: 168 registers in FP64 (12 warps), 128 in FP32 (16) */
#ifndef CLOUDSC_F64_WARPS_PER_SM
#define CLOUDSC_F64_WARPS_PER_SM 12   /* 16 = a 128-register FP64 build (spills; measured slower, profiles/README.md) */
#endif
template <typename T> constexpr int warps_per_sm() { return sizeof(T) == 4 ? 16 : CLOUDSC_F64_WARPS_PER_SM; }

__device__ __forceinline__ unsigned smid_of_cta() {
  unsigned r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}

template <typename T>
__global__ void __launch_bounds__(32, warps_per_sm<T>())
cloudsc_sweep_tmap(const __grid_constant__ FieldTable ft, const __grid_constant__ TensorMaps tm, const LaunchGeom g,
                   const __grid_constant__ StepConst<T> sc) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using IO = TmapIO<T>;
  const int ccol = g.col_begin + blockIdx.x * IO::kCols;   /* col_begin, nproma: multiples of 32 (launcher); col_end may be ragged */
  if (ccol >= g.col_end) return;
  const int ibl = ccol / g.nproma;
  const int jl0 = ccol - ibl * g.nproma;
  const int jl = jl0 + (int)threadIdx.x;
  const long long bL = (long long)ibl * g.klev * g.nproma, bH = (long long)ibl * (g.klev + 1) * g.nproma;
  IO io{ft, tm, ibl, jl0, g.nproma, g.klev, dev_consts<T>().ncldtop - 2,
        smem_u32(smem_raw), smem_u32(smem_raw + 2 * IO::kStage),
        reinterpret_cast<const T *>(smem_raw) + threadIdx.x, nullptr,
        bH + jl, (long long)ibl * g.nproma + jl,
        (unsigned)(bL + jl), (unsigned)(bH + jl + g.nproma), (unsigned)(bL * 5 + jl), (unsigned)(g.klev * g.nproma),
        ccol + (int)threadIdx.x < g.col_end, 0, g.klev, 0, nullptr, -1, 0};
  if (threadIdx.x == 0) {
    mbar_init(io.bar0, 1);
    mbar_init(io.bar0 + 8, 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (elect_one()) {
    io.issue(0);
    if (g.klev > 1) io.issue(1);
  }
  __syncwarp();
  /* -DCLOUDSC_TILE_TIMES (tools/gpu_tile_times.py): start / end (%globaltimer, ns) and SM of every tile into the TENDENCY_CML_T
     buffer, which the kernel never touches otherwise */
#ifdef CLOUDSC_TILE_TIMES
  long long tt0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt0));
#endif
  cloudsc_column<T>(dev_consts<T>(), sc, io, g.klev);
#ifdef CLOUDSC_TILE_TIMES
  if (threadIdx.x == 0 && ft.p[CLOUDSC_F_TENDENCY_CML_T]) {
    long long tt1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt1));
    long long *d = static_cast<long long *>(ft.p[CLOUDSC_F_TENDENCY_CML_T]) + 3 * (size_t)blockIdx.x;
    d[0] = tt0; d[1] = tt1; d[2] = (long long)smid_of_cta();
  }
#endif
}

/* ---- persistent kernel: every warp walks its own column tiles, upper and lower parts interleaved ------------------- */
/*
 * A sweep has two regimes: the levels above the clouds move ~50 rows of IO for a few hundred
 * instructions and run at the HBM peak once all resident warps are there; the cloudy levels are latency/issue-bound
 * while HBM idles at ~50 %.  With one CTA per tile all warps start at level 0 together, so the regimes never overlap.
 * Here the grid is exactly the resident set (12 one-warp CTAs per SM); warp w owns tiles w, w + G, w + 2G, ... and runs
 * each of them in two pieces, A = levels [0, ksplit) and B = [ksplit, klev), the state at the cut going through
 * `plan.carry`.  Whenever a warp may choose between the B piece of a tile it has started and the A piece of its next
 * tile it looks at how many warps of ITS SM are in a B piece right now (a per-SM counter) and takes B only below
 * `plan.target_b`: the SM then holds a steady mix of bandwidth-bound and latency-bound warps.  There is no dependency
 * between warps, only that counter; results are bit-identical to the one-CTA-per-tile kernel for any ksplit
 * (tests/test_cpu_oracle.py::test_cut_sweep_is_bit_identical, tests/test_gpu_parity.py).
 */
struct PersistPlan {
  int ksplit;      /* first level of the B piece, 1 <= ksplit < klev */
  int ntiles;      /* 32-column tiles of the launch */
  int target_b;    /* warps per SM that should be in a B piece */
  void *carry;     /* [ntiles][kCarryVals][32] elements */
  int *sm_b;       /* per-SM count of warps inside a B piece (self-resetting: +1 / -1 per piece) */
};

__device__ __forceinline__ unsigned smid() {
  unsigned r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}

template <typename T>
__global__ void __launch_bounds__(32, warps_per_sm<T>())
cloudsc_sweep_persist(const __grid_constant__ FieldTable ft, const __grid_constant__ TensorMaps tm, const LaunchGeom g,
                      const __grid_constant__ StepConst<T> sc, const PersistPlan plan) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using IO = TmapIO<T, true>;
  const int w = (int)blockIdx.x, G = (int)gridDim.x;
  if (w >= plan.ntiles) return;
  const int n_mine = (plan.ntiles - w + G - 1) / G;
  const unsigned bar0 = smem_u32(smem_raw + 2 * IO::kStage);
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    mbar_fence_init();
  }
  __syncwarp();
  int *const sm_b = plan.sm_b + smid();
  int a_done = 0, b_done = 0, staged = 0;   /* staged: levels this warp has put through the ring so far */
  while (b_done < n_mine) {
    /* which piece next (lane 0 decides, the warp follows) */
    int do_b = 0;
    if (threadIdx.x == 0) {
      const bool can_a = a_done < n_mine, can_b = b_done < a_done;
      if (!can_a) { do_b = 1; atomicAdd(sm_b, 1); }
      else if (can_b) {
        if (atomicAdd(sm_b, 1) < plan.target_b) do_b = 1;
        else atomicSub(sm_b, 1);
      }
    }
    do_b = __shfl_sync(0xffffffffu, do_b, 0);
    const int tile = w + (do_b ? b_done : a_done) * G;
    const int kb = do_b ? plan.ksplit : 0, ke = do_b ? g.klev : plan.ksplit;
    const int ccol = g.col_begin + tile * IO::kCols;
    const int ibl = ccol / g.nproma;
    const int jl0 = ccol - ibl * g.nproma;
    const int jl = jl0 + (int)threadIdx.x;
    const long long bL = (long long)ibl * g.klev * g.nproma, bH = (long long)ibl * (g.klev + 1) * g.nproma;
    const unsigned rows = (unsigned)kb * (unsigned)g.nproma;   /* output offsets start at the piece's first level */
    IO io{ft, tm, ibl, jl0, g.nproma, g.klev, dev_consts<T>().ncldtop - 2,
          smem_u32(smem_raw), bar0,
          reinterpret_cast<const T *>(smem_raw) + threadIdx.x, nullptr,
          bH + jl, (long long)ibl * g.nproma + jl,
          (unsigned)(bL + jl) + rows, (unsigned)(bH + jl + g.nproma) + rows, (unsigned)(bL * 5 + jl) + rows, (unsigned)(g.klev * g.nproma),
          ccol + (int)threadIdx.x < g.col_end, kb, ke, staged - kb,
          static_cast<T *>(plan.carry) + (size_t)tile * (kCarryVals * 32) + threadIdx.x, -1, 0};
    if (elect_one()) {
      io.issue(kb);
      if (kb + 1 < ke) io.issue(kb + 1);
    }
    __syncwarp();
    cloudsc_column<T>(dev_consts<T>(), sc, io, g.klev);
    __syncwarp();             /* all lanes are done with the ring and the scratch before the next piece re-uses them */
    staged += ke - kb;
    if (do_b) { ++b_done; if (threadIdx.x == 0) atomicSub(sm_b, 1); }
    else ++a_done;
  }
}

/* ---- streaming kernel: persistent warps, tiles from a queue, the input ring never drains --------------------------- */
/*
 * Schedule mode 2.  The grid is the resident set; every warp takes its next 32-column tile from a global counter (so the
 * tiles go to whichever warp is free, as with one CTA per tile) and sweeps it top to bottom in one piece, but requests the
 * first two levels of the NEXT tile during the last two levels of the current one: no CTA launch, no barrier set-up and no
 * empty input ring between two tiles of a warp.
 */
template <typename T>
__global__ void __launch_bounds__(32, warps_per_sm<T>())
cloudsc_sweep_stream(const __grid_constant__ FieldTable ft, const __grid_constant__ TensorMaps tm, const LaunchGeom g,
                     const __grid_constant__ StepConst<T> sc, const int ntiles, int *const queue) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using IO = TmapIO<T, true>;
  const unsigned bar0 = smem_u32(smem_raw + 2 * IO::kStage);
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    mbar_fence_init();
  }
  __syncwarp();
  /* the first gridDim.x tiles are dealt by block index, the rest through the queue (which starts at gridDim.x) */
  int tile = (int)blockIdx.x;
  if (tile >= ntiles) return;
  int staged = 0;
  bool primed = false;
  while (tile < ntiles) {
    int next = 0;
    if (threadIdx.x == 0) next = atomicAdd(queue, 1);
    next = __shfl_sync(0xffffffffu, next, 0);
    const int ccol = g.col_begin + tile * IO::kCols;
    const int ibl = ccol / g.nproma;
    const int jl0 = ccol - ibl * g.nproma;
    const int jl = jl0 + (int)threadIdx.x;
    const int ncol = g.col_begin + next * IO::kCols;
    const int n_ibl = next < ntiles ? ncol / g.nproma : -1;
    const int n_jl0 = next < ntiles ? ncol - n_ibl * g.nproma : 0;
    const long long bL = (long long)ibl * g.klev * g.nproma, bH = (long long)ibl * (g.klev + 1) * g.nproma;
    IO io{ft, tm, ibl, jl0, g.nproma, g.klev, dev_consts<T>().ncldtop - 2,
          smem_u32(smem_raw), bar0,
          reinterpret_cast<const T *>(smem_raw) + threadIdx.x, nullptr,
          bH + jl, (long long)ibl * g.nproma + jl,
          (unsigned)(bL + jl), (unsigned)(bH + jl + g.nproma), (unsigned)(bL * 5 + jl), (unsigned)(g.klev * g.nproma),
          ccol + (int)threadIdx.x < g.col_end, 0, g.klev, staged, nullptr, n_ibl, n_jl0};
    if (!primed) {
      if (elect_one()) {
        io.issue(0);
        io.issue(1);
      }
      __syncwarp();
      primed = true;
    }
    cloudsc_column<T>(dev_consts<T>(), sc, io, g.klev);
    __syncwarp();
    staged += g.klev;
    tile = next;
  }
}

/* ---- split-tail launch: the last slots of a launch are filled with half tiles ------------------------------------------ */
/*
 * Schedule mode 3.  With one CTA per tile the launch ends on whole tiles of ~0.55 ms: in the last 150 us only 1 149 -> 171 of
 * the 1 776 slots still hold one, 10 % of the launch (profiles/r2_tile_times.txt).  Here `nsplit` tiles are swept in two
 * pieces by different CTAs of ONE launch, in block-index order (the order the hardware dispatches them in):
 *   blocks [0, nsplit)                    levels [0, cut) of tile b; the state at the cut (`Carry` + parked slots, as mode 1)
 *                                         goes to `carry`, then flag[b] = seq
 *   blocks [nsplit, ntiles)               whole tiles
 *   blocks [ntiles, ntiles + nsplit)      levels [cut, klev) of tile b - ntiles, after flag[..] == seq (set about two tile
 *                                         times earlier: the wait is there for correctness, not for timing)
 * so the pieces that fill the end of the launch are about half a tile long.  `seq` changes with every launch: no flag is
 * ever reset.  Bit-identical to the other schedules of this instantiation (a cut sweep is: test_cut_sweep_is_bit_identical).
 */
struct SplitPlan {
  int ntiles, nsplit, cut, seq;
  void *carry;     /* [nsplit][kCarryVals][32] elements */
  int *flag;       /* [nsplit] */
};

/* The upper piece of a tile ran about two tile times before its lower piece is dispatched (blocks go out in index order); make
 * sure, and see its parked state.  EVERY lane polls: a wait under `lane == 0` is divergent control flow in front of the sweep
 * and costs it its uniform-datapath code (+350 instructions, 24 bytes of spills inside the level loop).  A launch whose upper
 * pieces never ran (it would have to be dispatched out of order) traps after ~2 s instead of hanging the device. */
__device__ __forceinline__ void split_wait(const int *flag, int seq) {
  for (int n = 0;; ++n) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v == seq) break;
    __nanosleep(500);
    if (n > (1 << 22)) __trap();
  }
}

template <typename T>
__global__ void __launch_bounds__(32, warps_per_sm<T>())
cloudsc_sweep_split(const __grid_constant__ FieldTable ft, const __grid_constant__ TensorMaps tm, const LaunchGeom g,
                    const __grid_constant__ StepConst<T> sc, const __grid_constant__ SplitPlan plan) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using IO = TmapIO<T, true>;
  /* the block plan (cloudsc_split_plan.h; declares b, tail, tile, split, kb, ke -- the text split_piece() expands for the host) */
  CLOUDSC_SPLIT_PIECE((int)blockIdx.x, plan.ntiles, plan.nsplit, plan.cut, g.klev);
  const int ccol = g.col_begin + tile * IO::kCols;
  if (ccol >= g.col_end) return;
  const int ibl = ccol / g.nproma;
  const int jl0 = ccol - ibl * g.nproma;
  const int jl = jl0 + (int)threadIdx.x;
  const long long bL = (long long)ibl * g.klev * g.nproma, bH = (long long)ibl * (g.klev + 1) * g.nproma;
  const unsigned rows = (unsigned)kb * (unsigned)g.nproma;   /* output offsets start at the piece's first level */
  const unsigned bar0 = smem_u32(smem_raw + 2 * IO::kStage);
  IO io{ft, tm, ibl, jl0, g.nproma, g.klev, dev_consts<T>().ncldtop - 2,
        smem_u32(smem_raw), bar0,
        reinterpret_cast<const T *>(smem_raw) + threadIdx.x, nullptr,
        bH + jl, (long long)ibl * g.nproma + jl,
        (unsigned)(bL + jl) + rows, (unsigned)(bH + jl + g.nproma) + rows, (unsigned)(bL * 5 + jl) + rows, (unsigned)(g.klev * g.nproma),
        ccol + (int)threadIdx.x < g.col_end, kb, ke, -kb,
        static_cast<T *>(plan.carry) + (size_t)(split ? tile : 0) * (kCarryVals * 32) + threadIdx.x, -1, 0};
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    mbar_fence_init();
  }
  if (tail) split_wait(&plan.flag[tile], plan.seq);   /* every lane polls: a wait under `lane == 0` is divergent control flow */
  __syncwarp();
  if (elect_one()) {
    io.issue(kb);
    if (kb + 1 < ke) io.issue(kb + 1);
  }
  __syncwarp();
  cloudsc_column<T>(dev_consts<T>(), sc, io, g.klev);
  if (io.ke < g.klev) {         /* an upper piece: blocks [0, nsplit), tile == block index */
    __threadfence();            /* every lane: its part of the parked state is visible before the flag */
    __syncwarp();
    if (threadIdx.x == 0) atomicExch(&plan.flag[blockIdx.x], plan.seq);
  }
}

/* ---- IO policy 3: thread-private cp.async ring, pre-offset slot pointers, streaming stores -------- */
/*
 * Every thread stages ITS OWN column: per level 27 `cp.async` (SASS LDGSTS, 8 bytes each, a warp
 * instruction moves one coalesced 256-byte row) copy the inputs of level k+2 into the thread's slots
 * of a 2-stage shared-memory ring while level k is computed; `cp.async.wait_group 1` at the top of a
 * level is the only synchronisation -- no barrier, no mbarrier, no coupling between threads, any
 * NPROMA / column range.  The consumer reads a slot with one LDS at an immediate offset.
 *
 * Address arithmetic is reduced to one 64-bit add per access: the host folds everything that does
 * not depend on the thread (species offset inside [nclv][klev][nproma] fields, the +1 row of the
 * fields consumed one level ahead / of the [klev+1] flux outputs) into a table of 27 + 24 slot
 * pointers passed in the kernel-parameter bank, and the thread keeps three running byte offsets
 * (one per field kind) that it bumps by one row per level.
 */
struct SlotTable {
  const void *in[N_IN];   /* (field, species, row shift) of every input slot */
  void *out[N_OUT];       /* same for the output slots */
  const void *paph, *plsm;
  const int *ktype;
  void *rainfrac;
};

/* 1 for the output slots written at row k+1 (the [klev+1] flux fields) */
__host__ __device__ constexpr int out_next(int s) { return s >= OUT_FSQLF ? 1 : 0; }

template <typename T> __device__ __forceinline__ void cp_async_elem(unsigned dst, const void *src);
template <> __device__ __forceinline__ void cp_async_elem<double>(unsigned dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
template <> __device__ __forceinline__ void cp_async_elem<float>(unsigned dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

template <typename T, int BLOCK>
struct PipeIO {
  static constexpr int kRow = BLOCK * (int)sizeof(T);      /* bytes between two slots of a thread */
  static constexpr int kStage = N_IN * kRow;
  static constexpr int kSmem = 2 * kStage;
  const SlotTable &tab;
  long long bL, bH, bC;     /* byte offsets of (block, current level, this column) per field kind */
  long long b1;             /* element offset of the per-column scalars */
  long long rowb;           /* nproma * sizeof(T) */
  int klev;
  unsigned sm_t;            /* shared address of this thread's slot 0 of stage 0 */
  const T *st_t;            /* generic pointer to the same */
  const T *cur;             /* this thread's slot 0 of the stage being consumed */

  /* queue the inputs of `level`; (pL, pH, pC) = byte offsets of that level */
  __device__ __forceinline__ void issue(int level, long long pL, long long pH, long long pC) {
    if (level < klev) {
      const unsigned dst = sm_t + ((unsigned)level & 1u) * (unsigned)kStage;
      const bool has_next = level + 1 < klev;   /* pmfu/pmfd/plu have no row klev; it is never consumed */
#pragma unroll
      for (int s = 0; s < N_IN; ++s) {
        const long long off = in_kind(s) == 0 ? pL : in_kind(s) == 1 ? pH : pC;
        if (s == IN_PAPH || !in_next(s) || has_next)
          cp_async_elem<T>(dst + (unsigned)(s * kRow), static_cast<const char *>(tab.in[s]) + off);
      }
    }
    cp_async_commit();   /* one group per level, empty past the last one: keeps wait_group 1 exact */
  }
  __device__ __forceinline__ void prologue() {
    issue(0, bL, bH, bC);
    issue(1, bL + rowb, bH + rowb, bC + rowb);
  }
  __device__ __forceinline__ void level_begin(int k) {
    cur = st_t + (k & 1) * (kStage / (int)sizeof(T));
    cp_async_wait1();
  }
  __device__ __forceinline__ void level_end(int k) {
    issue(k + 2, bL + 2 * rowb, bH + 2 * rowb, bC + 2 * rowb);
    bL += rowb; bH += rowb; bC += rowb;
  }
  template <int S> __device__ __forceinline__ T ld(int) const { return cur[S * BLOCK]; }
  template <int S> __device__ __forceinline__ T ld_next(int) const { return cur[S * BLOCK]; }
  __device__ __forceinline__ T paph_row(int r) const {
    return __ldg(reinterpret_cast<const T *>(static_cast<const char *>(tab.paph) + bH + (long long)r * rowb));
  }
  template <int S> __device__ __forceinline__ void st(int, T v) {
    const long long off = out_kind(S) == 0 ? bL : out_kind(S) == 1 ? bH : bC;
    __stcs(reinterpret_cast<T *>(static_cast<char *>(tab.out[S]) + off), v);
  }
  /* row 0 of a flux field: tab.out carries the +1 row of the in-loop stores */
  template <int S> __device__ __forceinline__ void st_row0(T v) {
    __stcs(reinterpret_cast<T *>(static_cast<char *>(tab.out[S]) + (bH - rowb)), v);
  }
  __device__ __forceinline__ T plsm() const { return __ldg(static_cast<const T *>(tab.plsm) + b1); }
  __device__ __forceinline__ int ktype() const { return __ldg(tab.ktype + b1); }
  __device__ __forceinline__ void st_rainfrac(T v) { static_cast<T *>(tab.rainfrac)[b1] = v; }
};

template <typename T, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
cloudsc_sweep_pipe(const __grid_constant__ SlotTable tab, const LaunchGeom g, const __grid_constant__ StepConst<T> sc) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using IO = PipeIO<T, BLOCK>;
  const int col = g.col_begin + blockIdx.x * BLOCK + threadIdx.x;
  if (col >= g.col_end) return;
  const int ibl = col / g.nproma;
  const int jl = col - ibl * g.nproma;
  const long long eL = (long long)ibl * g.klev * g.nproma;
  IO io{tab,
        (eL + jl) * (long long)sizeof(T),
        ((long long)ibl * (g.klev + 1) * g.nproma + jl) * (long long)sizeof(T),
        (eL * 5 + jl) * (long long)sizeof(T),
        (long long)ibl * g.nproma + jl,
        (long long)g.nproma * (long long)sizeof(T), g.klev,
        smem_u32(smem_raw) + threadIdx.x * (unsigned)sizeof(T),
        reinterpret_cast<const T *>(smem_raw) + threadIdx.x, nullptr};
  io.prologue();
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

/* launches per kernel family since the library was loaded (cloudsc_b200_launch_counts) */
static std::atomic<long long> g_launches[KERNEL_FAMILIES];
void launch_counts(long long out[KERNEL_FAMILIES]) {
  for (int i = 0; i < KERNEL_FAMILIES; ++i) out[i] = g_launches[i].load();
}

namespace {
constexpr int kDirectBlock = 128;

/* Every kernel is built for ONE register budget: 168 registers = 3 warps per scheduler = 12 per SM.  255 registers
   (8 warps/SM) lose more latency hiding than they gain in scheduling freedom, 128 registers (16 warps/SM) spill
   and still need 3 waves for 163 840 columns (profiles/README.md); those builds were measured and dropped. */


template <typename T, bool AER>
cudaError_t launch_direct(const FieldTable &ft, const LaunchGeom &g, double ptsphy, cudaStream_t stream) {
  const int ncols = g.col_end - g.col_begin;
  const int grid = (ncols + kDirectBlock - 1) / kDirectBlock;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  StepConst<T> sc;
  fill_step<T>(host_consts<T>(dev), ptsphy, &sc);
  cloudsc_sweep_direct<T, kDirectBlock, 3, AER><<<grid, kDirectBlock, 0, stream>>>(ft, g, sc);
  ++g_launches[AER ? KERNEL_DIRECT_AER : KERNEL_DIRECT];
  return cudaGetLastError();
}

template <typename T>
const void *direct_fn(bool aer) {
  return aer ? (const void *)cloudsc_sweep_direct<T, kDirectBlock, 3, true> : (const void *)cloudsc_sweep_direct<T, kDirectBlock, 3, false>;
}

/* ---- tile (TMA-staged) variant ------------------------------------------------------------------ */
/* (warps per CTA, resident warps per SM the register budget is sized for: 8 -> 255 regs, 12 -> 168, 16 -> 128) */
struct TileCfg { int warps, wps; };
inline bool is_tile(int variant) { return variant == VARIANT_AUTO || variant == VARIANT_TILE_C4_W12 || variant == VARIANT_TILE_C1_W12 || variant == VARIANT_TILE_C2_W12; }
/* widest CTA tile whose 32*WARPS columns never straddle an NPROMA block */
inline int tile_warps_for(int nproma) { return nproma % 128 == 0 ? 4 : nproma % 64 == 0 ? 2 : nproma % 32 == 0 ? 1 : 0; }
inline TileCfg tile_cfg(int variant, int nproma) {
  const int widest = tile_warps_for(nproma);
  /* default: one-warp CTAs (warp-private tiles, no coupling between warps), 12 resident warps per SM
     -- the fastest configuration measured on B200 (profiles/README.md) */
  TileCfg c = {widest ? 1 : 0, 12};
  switch (variant) {
    case VARIANT_TILE_C4_W12: c.warps = widest; break;
    case VARIANT_TILE_C2_W12: c.warps = widest >= 2 ? 2 : widest; break;
    default: break;
  }
  if (c.warps != 4) c.wps = 12;   /* the narrow tiles are built for one register budget only */
  return c;
}

#define CLOUDSC_TILE_DISPATCH(cfg, ...)                                                   \
  if (cfg.warps == 4) { constexpr int W = 4, MB = 3; __VA_ARGS__; }                       \
  else if (cfg.warps == 2) { constexpr int W = 2, MB = 6; __VA_ARGS__; }                  \
  else { constexpr int W = 1, MB = warps_per_sm<T>(); __VA_ARGS__; }

template <typename T>
const void *tile_fn(TileCfg cfg, size_t *smem) {
  const void *fn = nullptr;
  CLOUDSC_TILE_DISPATCH(cfg, (fn = (const void *)cloudsc_sweep_tile<T, W, MB>, *smem = TileIO<T, W>::kSmem))
  return fn;
}

/* the TMA rows must be 16-byte aligned and a CTA's columns must not straddle NPROMA blocks */
template <typename T>
bool tile_ok(const FieldTable &ft, const LaunchGeom &g, TileCfg cfg) {
  if (cfg.warps == 0 || g.col_begin % (32 * cfg.warps) != 0) return false;
  /* the kernel addresses output rows with 32-bit element offsets */
  const long long nblk = (g.col_end + g.nproma - 1) / g.nproma;
  if (nblk * 5 * (g.klev + 1) * g.nproma >= (1ll << 32)) return false;
  for (int s = 0; s < N_IN; ++s)
    if (reinterpret_cast<uintptr_t>(ft.p[in_field(s)]) % 16 != 0) return false;
  return true;
}

template <typename T>
cudaError_t launch_tile(const FieldTable &ft, const LaunchGeom &g, double ptsphy, TileCfg cfg, cudaStream_t stream) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  size_t smem = 0;
  const void *fn = tile_fn<T>(cfg, &smem);
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  const int ncols = g.col_end - g.col_begin;
  const int per_cta = cfg.warps * 32;
  const int grid = (ncols + per_cta - 1) / per_cta;
  StepConst<T> sc;
  fill_step<T>(host_consts<T>(dev), ptsphy, &sc);
  void *args[] = {(void *)&ft, (void *)&g, (void *)&sc};
  ++g_launches[KERNEL_TILE];
  return cudaLaunchKernel(fn, dim3(grid), dim3(per_cta), args, smem, stream);
}

/* ---- pipe (cp.async) variant ---------------------------------------------------------------------- */
constexpr int kPipeBlock = 32;
inline bool is_pipe(int variant) { return variant == VARIANT_PIPE_W12; }

template <typename T>
SlotTable make_slot_table(const FieldTable &ft, const LaunchGeom &g) {
  SlotTable t;
  const long long row = g.nproma, sp = (long long)g.klev * g.nproma;
  for (int s = 0; s < N_IN; ++s)
    t.in[s] = static_cast<const char *>(ft.p[in_field(s)]) + (in_species(s) * sp + in_next(s) * row) * (long long)sizeof(T);
  for (int s = 0; s < N_OUT; ++s)
    t.out[s] = static_cast<char *>(ft.p[out_field(s)]) + (out_species(s) * sp + out_next(s) * row) * (long long)sizeof(T);
  t.paph = ft.p[CLOUDSC_F_PAPH];
  t.plsm = ft.p[CLOUDSC_F_PLSM];
  t.ktype = static_cast<const int *>(ft.p[CLOUDSC_F_KTYPE]);
  t.rainfrac = ft.p[CLOUDSC_F_PRAINFRAC_TOPRFZ];
  return t;
}

template <typename T>
const void *pipe_fn() {
  return (const void *)cloudsc_sweep_pipe<T, kPipeBlock, 12>;
}

template <typename T>
cudaError_t launch_pipe(const FieldTable &ft, const LaunchGeom &g, double ptsphy, cudaStream_t stream) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const void *fn = pipe_fn<T>();
  const size_t smem = PipeIO<T, kPipeBlock>::kSmem;
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  const int ncols = g.col_end - g.col_begin;
  const int grid = (ncols + kPipeBlock - 1) / kPipeBlock;
  SlotTable tab = make_slot_table<T>(ft, g);
  StepConst<T> sc;
  fill_step<T>(host_consts<T>(dev), ptsphy, &sc);
  void *args[] = {(void *)&tab, (void *)&g, (void *)&sc};
  ++g_launches[KERNEL_PIPE];
  return cudaLaunchKernel(fn, dim3(grid), dim3(kPipeBlock), args, smem, stream);
}
}  // namespace

template <typename T>
static cudaError_t launch_any(const FieldTable &ft, const LaunchGeom &g, double ptsphy, int variant, cudaStream_t stream) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  /* aerosol switches set (laericesed / laericeauto): the AER instantiation of the column code, plain loads */
  const CloudscConst<T> &hc = host_consts<T>(dev);
  if (hc.laericesed || hc.laericeauto) return launch_direct<T, true>(ft, g, ptsphy, stream);
  if (is_pipe(variant)) return launch_pipe<T>(ft, g, ptsphy, stream);
  const TileCfg cfg = tile_cfg(variant, g.nproma);
  if (is_tile(variant) && tile_ok<T>(ft, g, cfg)) {
    /* the TMA kernel moves whole rows of 32*WARPS columns; a ragged last tile reads the padding columns of the last
       NPROMA block (the fields are [nblocks][..][nproma], so they exist) and masks its stores */
    return launch_tile<T>(ft, g, ptsphy, cfg, stream);
  }
  /* geometry the TMA path cannot take (NPROMA not a multiple of 32, unaligned column range or
     pointers): same arithmetic, plain global loads */
  return launch_direct<T, false>(ft, g, ptsphy, stream);
}

template <typename T>
static cudaError_t launch_sweep_tmap_t(const FieldTable &ft, const TensorMaps &tm, const LaunchGeom &g, double ptsphy,
                                       cudaStream_t stream) {
  if (g.col_end <= g.col_begin) return cudaSuccess;
  /* the kernel addresses output rows with 32-bit element offsets (as the bulk-copy kernels, tile_ok()) */
  const long long nblk = (g.col_end + g.nproma - 1) / g.nproma;
  if (nblk * 5 * (g.klev + 1) * g.nproma >= (1ll << 32)) return launch_any<T>(ft, g, ptsphy, VARIANT_AUTO, stream);
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (host_consts<T>(dev).laericesed || host_consts<T>(dev).laericeauto) return launch_any<T>(ft, g, ptsphy, VARIANT_AUTO, stream);
  const void *fn = (const void *)cloudsc_sweep_tmap<T>;
  const size_t smem = TmapIO<T>::kSmem;
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  /* whole 32-column rows through TMA; the lanes of a ragged last tile beyond col_end mask their stores */
  StepConst<T> sc;
  fill_step<T>(host_consts<T>(dev), ptsphy, &sc);
  const int grid = (g.col_end - g.col_begin + 31) / 32;
  void *args[] = {(void *)&ft, (void *)&tm, (void *)&g, (void *)&sc};
  ++g_launches[KERNEL_TMAP];
  return cudaLaunchKernel(fn, dim3(grid), dim3(32), args, smem, stream);
}

cudaError_t launch_sweep_tmap(const FieldTable &ft, const TensorMaps &tm, const LaunchGeom &g, double ptsphy,
                              int precision, cudaStream_t stream) {
  if (precision == CLOUDSC_B200_FP32) return launch_sweep_tmap_t<float>(ft, tm, g, ptsphy, stream);
  return launch_sweep_tmap_t<double>(ft, tm, g, ptsphy, stream);
}

/* Which kernel a schedule gets for one launch: its own (returns ps.mode) or one CTA per tile (returns 0) -- aerosol runs,
 * launches too small for the schedule (below) and states without scratch keep the latter.
 * sms_out / per_out: SM count and resident CTAs per SM of the persistent kernels. */
template <typename T>
static cudaError_t persist_decide(const LaunchGeom &g, const PersistSchedule &ps, int *mode_out, int *sms_out, int *per_out) {
  *mode_out = 0;
  const long long nblk = (g.col_end + g.nproma - 1) / g.nproma;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const int ntiles = (g.col_end - g.col_begin + 31) / 32;
  const bool aer = host_consts<T>(dev).laericesed || host_consts<T>(dev).laericeauto;
  if (ps.mode < 1 || ps.mode > 3 || aer || nblk * 5 * (g.klev + 1) * g.nproma >= (1ll << 32) || g.klev < 3 || !ps.carry || !ps.sm_b ||
      ntiles > ps.carry_tiles)
    return cudaSuccess;
  const void *fn = (const void *)cloudsc_sweep_persist<T>;
  const size_t smem = TmapIO<T, true>::kSmem;
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  static int sms[64] = {0}, per_sm[64][2] = {{0}};
  const int pi = sizeof(T) == 4 ? 1 : 0;
  if (!sms[dev & 63]) {
    e = cudaDeviceGetAttribute(&sms[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
  }
  if (!per_sm[dev & 63][pi]) {
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[dev & 63][pi], fn, 32, smem);
    if (e != cudaSuccess) return e;
  }
  *sms_out = sms[dev & 63];
  *per_out = per_sm[dev & 63][pi];
  /* The persistent kernels (modes 1, 2) need two full rounds of the resident set to have anything to interleave.  The
     split-tail launch pays as soon as there is a second round at all -- 1 792 tiles (1.01 rounds) x1.19, 2 048 x1.15,
     2 816 x1.06, 3 752 x1.15, 5 120 x1.05, 20 480 x1.02 -- and costs 2-3 % when every tile is resident at once (1 024 tiles:
     profiles/r2_split_small.txt).  Measurement knob: CLOUDSC_B200_SPLIT_MIN_TILES=<n> moves its bar. */
  const int resident = *sms_out * *per_out;
  int min_tiles = 2 * resident;
  if (ps.mode == 3) {
    min_tiles = resident + 1;
    static const int env_min = [] { const char *v = getenv("CLOUDSC_B200_SPLIT_MIN_TILES"); return v ? atoi(v) : 0; }();
    if (env_min > 0) min_tiles = env_min;
  }
  if (resident > 0 && ntiles >= min_tiles) *mode_out = ps.mode;
  return cudaSuccess;
}

/* persistent kernels: grid = the resident set (SMs x warps per SM), never more than the tiles */
template <typename T>
static cudaError_t launch_sweep_persist_t(const FieldTable &ft, const TensorMaps &tm, const LaunchGeom &g, double ptsphy,
                                          const PersistSchedule &ps, cudaStream_t stream) {
  if (g.col_end <= g.col_begin) return cudaSuccess;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  int use = 0, nsm = 0, per_cta = 0;
  e = persist_decide<T>(g, ps, &use, &nsm, &per_cta);
  if (e != cudaSuccess) return e;
  if (use == 0) return launch_sweep_tmap_t<T>(ft, tm, g, ptsphy, stream);
  const int ntiles = (g.col_end - g.col_begin + 31) / 32;
  const void *fn = (const void *)cloudsc_sweep_persist<T>;
  const size_t smem = TmapIO<T, true>::kSmem;
  const int resident = nsm * per_cta;
  static int resident_h[64];                          /* stable host storage for the asynchronous queue reset */
  resident_h[dev & 63] = resident;
  if (ps.mode == 2) {
    const void *sfn = (const void *)cloudsc_sweep_stream<T>;
    e = cudaFuncSetAttribute(sfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(sfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    int *queue = ps.sm_b + 512;                       /* (the per-SM counters of mode 1 use the first half) */
    e = cudaMemcpyAsync(queue, &resident_h[dev & 63], sizeof(int), cudaMemcpyHostToDevice, stream);   /* queue starts at the grid size */
    if (e != cudaSuccess) return e;
    StepConst<T> sc2;
    fill_step<T>(host_consts<T>(dev), ptsphy, &sc2);
    int nt = ntiles;
    void *sargs[] = {(void *)&ft, (void *)&tm, (void *)&g, (void *)&sc2, (void *)&nt, (void *)&queue};
    ++g_launches[KERNEL_TMAP];
    return cudaLaunchKernel(sfn, dim3(resident), dim3(32), sargs, smem, stream);
  }
  const int ksplit = split_cut(ps.ksplit, g.klev);
  if (ps.mode == 3) {
    const void *sfn = (const void *)cloudsc_sweep_split<T>;
    e = cudaFuncSetAttribute(sfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(sfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    static std::atomic<int> seq{0};   /* one value per launch, process-wide: two host threads never hand out the same one */
    SplitPlan sp;
    sp.ntiles = ntiles;
    sp.nsplit = split_count(ntiles, nsm, per_cta, ps.target_b, kSplitFlags, ps.carry_tiles);   /* 5 of 12 tiles per SM (profiles/r2_split_tail.txt) */
    sp.cut = ksplit;
    sp.seq = ++seq;
    if (sp.seq == 0) sp.seq = ++seq;    /* 0 is what a fresh flag holds */
    sp.carry = ps.carry;
    sp.flag = ps.sm_b + kSplitFlagOffset;
    StepConst<T> sc3;
    fill_step<T>(host_consts<T>(dev), ptsphy, &sc3);
    void *sargs[] = {(void *)&ft, (void *)&tm, (void *)&g, (void *)&sc3, (void *)&sp};
    ++g_launches[KERNEL_TMAP];
    return cudaLaunchKernel(sfn, dim3(split_blocks(ntiles, sp.nsplit)), dim3(32), sargs, smem, stream);
  }
  PersistPlan plan{ksplit, ntiles, ps.target_b > 0 ? ps.target_b : (per_cta * 7 + 6) / 12, ps.carry, ps.sm_b};
  StepConst<T> sc;
  fill_step<T>(host_consts<T>(dev), ptsphy, &sc);
  void *args[] = {(void *)&ft, (void *)&tm, (void *)&g, (void *)&sc, (void *)&plan};
  ++g_launches[KERNEL_TMAP];
  return cudaLaunchKernel(fn, dim3(resident), dim3(32), args, smem, stream);
}

cudaError_t launch_sweep_persist(const FieldTable &ft, const TensorMaps &tm, const LaunchGeom &g, double ptsphy, int precision,
                                 const PersistSchedule &ps, cudaStream_t stream) {
  if (precision == CLOUDSC_B200_FP32) return launch_sweep_persist_t<float>(ft, tm, g, ptsphy, ps, stream);
  return launch_sweep_persist_t<double>(ft, tm, g, ptsphy, ps, stream);
}
cudaError_t persist_schedule_used(const LaunchGeom &g, int precision, const PersistSchedule &ps, int *mode_out) {
  int nsm = 0, per_cta = 0;
  if (precision == CLOUDSC_B200_FP32) return persist_decide<float>(g, ps, mode_out, &nsm, &per_cta);
  return persist_decide<double>(g, ps, mode_out, &nsm, &per_cta);
}
size_t persist_carry_bytes(int ntiles, int precision) {
  return (size_t)ntiles * kCarryVals * 32 * (precision == CLOUDSC_B200_FP32 ? sizeof(float) : sizeof(double));
}

static cudaError_t tile_kernel_info(const void *fn, size_t dyn, int out[8]) {
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  cudaFuncAttributes a;
  e = cudaFuncGetAttributes(&a, fn);
  if (e != cudaSuccess) return e;
  int nb = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, 32, dyn);
  if (e != cudaSuccess) return e;
  out[0] = a.numRegs; out[1] = (int)a.localSizeBytes; out[2] = (int)a.sharedSizeBytes; out[3] = a.maxThreadsPerBlock;
  out[4] = nb; out[5] = 32; out[6] = 100; out[7] = (int)dyn;
  return cudaSuccess;
}

template <typename T>
static cudaError_t tmap_kernel_info_t(int mode, int out[8]) {
  /* out[6]: 100 = one CTA per tile, 100 + mode = the kernel of that schedule */
  const void *fn = mode == 1 ? (const void *)cloudsc_sweep_persist<T> : mode == 2 ? (const void *)cloudsc_sweep_stream<T>
                 : mode == 3 ? (const void *)cloudsc_sweep_split<T> : (const void *)cloudsc_sweep_tmap<T>;
  cudaError_t e = tile_kernel_info(fn, mode >= 1 && mode <= 3 ? TmapIO<T, true>::kSmem : TmapIO<T>::kSmem, out);
  if (e == cudaSuccess && mode >= 1 && mode <= 3) out[6] = 100 + mode;
  return e;
}

cudaError_t tmap_kernel_info(int precision, int out[8], int mode) {
  return precision == CLOUDSC_B200_FP32 ? tmap_kernel_info_t<float>(mode, out) : tmap_kernel_info_t<double>(mode, out);
}

/* ---- self-test of the fast-mode primitives (cloudsc_b200_selftest_primitives) ---------------------------------------- */
__global__ void selftest_primitives_kernel(const double *x, int n, double *out) {
  exp2_tab_init();
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n) return;
  const double v = x[i];
  const double zero = 0.0;
  (void)zero;
  typedef double T;
  double *o = out + 12 * (size_t)i;
  o[0] = CLOUDSC_MX0(v);
  o[1] = CLOUDSC_MN1(v);
  o[2] = CLOUDSC_MNH(v);
  o[3] = CLOUDSC_MX160(v);
  /* the clamp of exp_fast_lo, observable through exp: exp(-700) for everything below */
  o[4] = exp_fast_lo(v > 700.0 ? 700.0 : v);
  o[5] = exp_fast(v > 700.0 ? 700.0 : v < -700.0 ? -700.0 : v);
  o[6] = v != 0.0 ? log_fast(v < 0.0 ? -v : v) : 0.0;
  o[7] = v != 0.0 ? rcp_fast(v) : 0.0;
  o[8] = v != 0.0 ? sqrt_fast(v < 0.0 ? -v : v) : 0.0;
  o[9] = v != 0.0 ? div_fast(1.2345678901234567, v) : 0.0;
  {
    const double av = v < 0.0 ? -v : v;
    o[10] = pow_fast(av, 0.666);
    o[11] = pow_fast(av, 2.47);
  }
}
cudaError_t launch_selftest_primitives(const double *x, int n, double *out, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  selftest_primitives_kernel<<<(n + 127) / 128, 128, 0, stream>>>(x, n, out);
  return cudaGetLastError();
}

cudaError_t launch_sweep(const FieldTable &ft, const LaunchGeom &g, double ptsphy, int precision, int variant,
                         cudaStream_t stream) {
  if (g.col_end <= g.col_begin) return cudaSuccess;
  if (precision == CLOUDSC_B200_FP32) return launch_any<float>(ft, g, ptsphy, variant, stream);
  return launch_any<double>(ft, g, ptsphy, variant, stream);
}

cudaError_t sweep_kernel_info(int nproma, int precision, int variant, int out[8]) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const bool aer = h_const64[dev & 63].laericesed || h_const64[dev & 63].laericeauto;
  const TileCfg cfg = tile_cfg(variant, nproma);
  const bool pipe = !aer && is_pipe(variant);
  const bool tile = pipe || (!aer && is_tile(variant) && cfg.warps > 0);   /* "uses dynamic shared memory" */
  cudaFuncAttributes a;
  size_t dyn = 0;
  const void *fn;
  if (pipe) {
    fn = precision == CLOUDSC_B200_FP32 ? pipe_fn<float>() : pipe_fn<double>();
    dyn = precision == CLOUDSC_B200_FP32 ? PipeIO<float, kPipeBlock>::kSmem : PipeIO<double, kPipeBlock>::kSmem;
  } else {
    fn = precision == CLOUDSC_B200_FP32 ? (tile ? tile_fn<float>(cfg, &dyn) : direct_fn<float>(aer))
                                        : (tile ? tile_fn<double>(cfg, &dyn) : direct_fn<double>(aer));
  }
  if (tile) {
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
  }
  e = cudaFuncGetAttributes(&a, fn);
  if (e != cudaSuccess) return e;
  const int threads = pipe ? kPipeBlock : tile ? cfg.warps * 32 : kDirectBlock;
  int nb = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, threads, dyn);
  if (e != cudaSuccess) return e;
  out[0] = a.numRegs;
  out[1] = (int)a.localSizeBytes;
  out[2] = (int)a.sharedSizeBytes;
  out[3] = a.maxThreadsPerBlock;
  out[4] = nb;
  out[5] = threads;
  out[6] = variant;
  out[7] = (int)dyn;
  return cudaSuccess;
}

/* host mirror of the parameters set_params() uploaded to the current device */
int sweep_ncldtop() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  return h_const64[dev & 63].ncldtop;
}
int sweep_aerosol_mask() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  return (h_const64[dev & 63].laericesed ? 1 : 0) | (h_const64[dev & 63].laericeauto ? 2 : 0);
}

}  // namespace cloudsc
