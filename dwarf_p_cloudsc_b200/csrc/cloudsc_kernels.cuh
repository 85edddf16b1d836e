/*
 * cloudsc_kernels.cuh -- launch interface between the C ABI (cloudsc_abi.cu) and the sm_100a
 * kernels (cloudsc_kernels.cu).  Internal; not part of the public ABI.
 */
#ifndef CLOUDSC_KERNELS_CUH
#define CLOUDSC_KERNELS_CUH

#include <cuda.h>
#include <cuda_runtime.h>
#include "cloudsc_b200.h"
#include "cloudsc_consts.h"

namespace cloudsc {

/* the 56 field pointers of cloudsc_c(), passed to the kernel by value (448 B of param space) */
struct FieldTable {
  void *p[CLOUDSC_NFIELDS];
};

/* geometry of one launch: columns [col_begin, col_end) of a field set blocked by nproma */
struct LaunchGeom {
  int col_begin;  /* first global column (0-based) */
  int col_end;    /* one past the last global column */
  int nproma;
  int klev;
};

/* (the ids are stable across rounds: 1, 3, 4, 6, 10, 11 were aliases of 2, 5, 9 and have been removed) */
enum Variant {
  VARIANT_AUTO = 0,         /* = VARIANT_TILE_C1_W12 where NPROMA allows, else the direct kernel */
  /* one thread per column, loads/stores straight from/to global memory (any geometry; also the kernel of the
     aerosol parameter sets, laericesed / laericeauto) */
  VARIANT_DIRECT_R168 = 2,
  /* TMA-staged input tiles (cp.async.bulk + mbarrier) of 32*WARPS columns per CTA, streaming stores.
     Cn = warps per CTA (clamped to what NPROMA allows); all at 168 registers = 12 resident warps per SM */
  VARIANT_TILE_C4_W12 = 5,
  VARIANT_TILE_C1_W12 = 7,  /* one-warp CTAs, warp-private tiles */
  VARIANT_TILE_C2_W12 = 8,
  /* thread-private cp.async ring (LDGSTS), pre-offset slot pointers, streaming stores; any geometry */
  VARIANT_PIPE_W12 = 9,
  VARIANT_COUNT
};
inline bool variant_valid(int v) {
  return v == VARIANT_AUTO || v == VARIANT_DIRECT_R168 || v == VARIANT_TILE_C4_W12 || v == VARIANT_TILE_C1_W12 ||
         v == VARIANT_TILE_C2_W12 || v == VARIANT_PIPE_W12;
}

/*
 * Tensor maps of an ARENA-allocated input set (cloudsc_b200_create): fields of one kind lie at a
 * constant spacing, so ONE tensor-map TMA request fetches the rows of many fields.
 *   a : the 15 [klev] inputs consumed at their own level   box {32 cols, 1 level, 1 block, 15 fields}
 *   a_top : the same map with a 7-field box (what the levels above the physics read)
 *   b : plu, pmfu, pmfd (consumed one level ahead)          box {32, 1, 1, 3}
 *   c : pclv, tendency_tmp_cld, species 0..3                box {32, 1, 4 species, 1 block, 2 fields}
 *   h : paph                                                box {32, 1, 1}
 */
struct alignas(64) TensorMaps {
  CUtensorMap a, a_top, b, c, h;   /* a_top: the first kGroupATop = 7 fields of group a only (levels above the physics) */
};
/* order of the fields inside the arena groups (the kernel's shared-memory rows follow it) */
extern const int kArenaGroupA[15];
extern const int kArenaGroupB[3];
extern const int kArenaGroupC[2];

cudaError_t launch_sweep_tmap(const FieldTable &ft, const TensorMaps &tm, const LaunchGeom &g, double ptsphy,
                              int precision, cudaStream_t stream);
/* persistent variant of the tensor-map kernel (cloudsc_kernels.cu, "persistent kernel"); falls back to launch_sweep_tmap
 * for launches too small for the schedule (modes 1, 2: fewer than two rounds of tiles per resident warp; mode 3: every tile
 * resident at once).  ksplit / target_b <= 0: defaults. */
struct PersistSchedule {
  int mode;         /* 1: interleaved upper / lower pieces (cloudsc_sweep_persist); 2: streaming tile queue (cloudsc_sweep_stream);
                       3: split tail (cloudsc_sweep_split; ksplit = cut level, target_b = split tiles per SM) */
  int ksplit, target_b;
  void *carry;      /* persist_carry_bytes(carry_tiles, precision) bytes of device scratch */
  int carry_tiles;
  int *sm_b;        /* >= kSchedInts zero-initialised ints of device memory */
};
/* layout of PersistSchedule::sm_b: [0, 512) per-SM counters of mode 1, [512] the tile queue of mode 2, [kSplitFlagOffset, +kSplitFlags)
 * the "upper piece done" flags of mode 3 */
constexpr int kSplitFlagOffset = 1024, kSplitFlags = 4096, kSchedInts = kSplitFlagOffset + kSplitFlags;
cudaError_t launch_sweep_persist(const FieldTable &ft, const TensorMaps &tm, const LaunchGeom &g, double ptsphy, int precision,
                                 const PersistSchedule &ps, cudaStream_t stream);
size_t persist_carry_bytes(int ntiles, int precision);
/* the schedule launch_sweep_persist would use for this launch: ps.mode, or 0 = one CTA per tile (small launch, aerosol run ...) */
cudaError_t persist_schedule_used(const LaunchGeom &g, int precision, const PersistSchedule &ps, int *mode_out);
cudaError_t tmap_kernel_info(int precision, int out[8], int mode = 0);   /* mode: the schedule whose kernel is described */

cudaError_t launch_selftest_primitives(const double *x, int n, double *out, cudaStream_t stream);
cudaError_t upload_consts(const CloudscConst<double> &c64, const CloudscConst<float> &c32);
cudaError_t launch_sweep(const FieldTable &ft, const LaunchGeom &g, double ptsphy, int precision, int variant,
                         cudaStream_t stream);
cudaError_t sweep_kernel_info(int nproma, int precision, int variant, int out[8]);
/* kernel families, for the launch counters */
enum KernelFamily { KERNEL_TMAP, KERNEL_TILE, KERNEL_DIRECT, KERNEL_PIPE, KERNEL_DIRECT_AER, KERNEL_FAMILIES };
void launch_counts(long long out[KERNEL_FAMILIES]);
/* host mirrors of what set_params() uploaded to the current device: ncldtop; bit 0 = laericesed, bit 1 = laericeauto */
int sweep_ncldtop();
int sweep_aerosol_mask();

}  // namespace cloudsc
#endif
