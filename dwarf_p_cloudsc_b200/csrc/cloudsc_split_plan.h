/*
 * The block plan of the split-tail launch (schedule mode 3, cloudsc_kernels.cu: cloudsc_sweep_split) -- which tile and which
 * level range a block of the launch sweeps, and how many tiles a launch splits.  Plain C++ shared by the kernel, its
 * launcher and the host emulation of the schedule (tests/host_emul/emul.cpp), so that the CPU tests walk the very plan
 * the device runs.
 *
 *   blocks [0, nsplit)                    levels [0, cut)    of tile b            (upper piece: parks the state, sets the flag)
 *   blocks [nsplit, ntiles)               levels [0, klev)   of tile b            (whole tile)
 *   blocks [ntiles, ntiles + nsplit)      levels [cut, klev) of tile b - ntiles   (lower piece: waits for the flag)
 *
 * Blocks are dispatched in index order, so every upper piece has started before any lower piece is looked at.
 */
#ifndef CLOUDSC_SPLIT_PLAN_H
#define CLOUDSC_SPLIT_PLAN_H

#if defined(__CUDACC__)
#define CLOUDSC_PLAN_FN __host__ __device__ __forceinline__
#else
#define CLOUDSC_PLAN_FN static inline
#endif

namespace cloudsc {

struct SplitPiece {
  int tile;      /* 32-column tile of the launch */
  int kb, ke;    /* level range [kb, ke) */
  bool lower;    /* a lower piece: the state at `kb` comes from the tile's parked carry */
  bool upper;    /* an upper piece: the state at `ke` goes to the tile's parked carry */
};

/* The plan itself, as declarations of five locals.  A macro, not a call: the kernel expands it in place (routing the values
 * through a returned struct moves ~90 instructions of the sweep kernel around, and the binary that was measured is kept),
 * split_piece() below expands the same text for the host. */
#define CLOUDSC_SPLIT_PIECE(b_, ntiles_, nsplit_, cut_, klev_)                 \
  const int b = (b_);                                                          \
  const bool tail = b >= (ntiles_);                                            \
  const int tile = tail ? b - (ntiles_) : b;                                   \
  const bool split = tile < (nsplit_);                                         \
  const int kb = tail ? (cut_) : 0, ke = (split && !tail) ? (cut_) : (klev_)

CLOUDSC_PLAN_FN SplitPiece split_piece(int block, int ntiles, int nsplit, int cut, int klev) {
  CLOUDSC_SPLIT_PIECE(block, ntiles, nsplit, cut, klev);
  SplitPiece p;
  p.tile = tile;
  p.kb = kb;
  p.ke = ke;
  p.lower = tail;
  p.upper = split && !tail;
  return p;
}

/* grid size of the launch */
CLOUDSC_PLAN_FN int split_blocks(int ntiles, int nsplit) { return ntiles + nsplit; }

/* tiles a launch splits: `per_sm` tiles per SM (<= 0: 5 of every 12 resident CTAs), never more than the flags, the
 * scratch or the launch hold */
CLOUDSC_PLAN_FN int split_count(int ntiles, int nsm, int resident_per_sm, int per_sm, int max_flags, int carry_tiles) {
  int per = per_sm > 0 ? per_sm : (resident_per_sm * 5 + 6) / 12;
  if (per > resident_per_sm) per = resident_per_sm;
  int n = per * nsm;
  if (n > max_flags) n = max_flags;
  if (n > carry_tiles) n = carry_tiles;
  if (n > ntiles) n = ntiles;
  return n;
}

/* cut level: `ksplit` (<= 0: 3/5 of the levels), inside [1, klev - 1] */
CLOUDSC_PLAN_FN int split_cut(int ksplit, int klev) {
  int c = ksplit > 0 ? ksplit : (3 * klev) / 5;
  if (c < 1) c = 1;
  if (c > klev - 1) c = klev - 1;
  return c;
}

}  // namespace cloudsc
#endif
