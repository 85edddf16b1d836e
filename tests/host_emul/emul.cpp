/*
 * emul.cpp -- TEST-ONLY host build of the device column code (csrc/cloudsc_column.cuh).
 *
 * Compiles the exact header the CUDA kernels instantiate, for the host, with FP contraction off,
 * so that `-m "not gpu"` tests can check the re-organised (k-cached, sparse-solve) arithmetic
 * without a GPU.  Built twice by oracle/Makefile: -DCLOUDSC_STRICT=1 (libcloudsc_emul.so, must be
 * bit-identical to the reference kernel) and -DCLOUDSC_STRICT=0 (libcloudsc_emul_fast.so, the
 * arithmetic mode of the product kernels: shared reciprocals, polynomial exp).  Never linked into, loaded by, or
 * reachable from the product library: libcloudsc_b200.so has no host arithmetic path.
 */
#include <cstddef>
#include <cstring>
#include "cloudsc_b200.h"
#include "cloudsc_column.cuh"
#include <vector>
#include "cloudsc_fill.h"
#include "cloudsc_split_plan.h"

namespace {
using namespace cloudsc;

template <typename T>
struct HostIO {
  void *const *f;  /* block-offset field table */
  int jl, stride, klev;
  static int in_field(int s) {
    static const int m[N_IN] = {CLOUDSC_F_PT, CLOUDSC_F_PQ, CLOUDSC_F_TENDENCY_TMP_T, CLOUDSC_F_TENDENCY_TMP_Q,
      CLOUDSC_F_TENDENCY_TMP_A, CLOUDSC_F_PA, CLOUDSC_F_PAP, CLOUDSC_F_PSUPSAT, CLOUDSC_F_PLUDE, CLOUDSC_F_PLU,
      CLOUDSC_F_PSNDE, CLOUDSC_F_PMFU, CLOUDSC_F_PMFD, CLOUDSC_F_PHRSW, CLOUDSC_F_PHRLW, CLOUDSC_F_PVERVEL,
      CLOUDSC_F_PVFL, CLOUDSC_F_PVFI, CLOUDSC_F_PCLV, CLOUDSC_F_PCLV, CLOUDSC_F_PCLV, CLOUDSC_F_PCLV,
      CLOUDSC_F_TENDENCY_TMP_CLD, CLOUDSC_F_TENDENCY_TMP_CLD, CLOUDSC_F_TENDENCY_TMP_CLD, CLOUDSC_F_TENDENCY_TMP_CLD,
      CLOUDSC_F_PAPH};
    return m[s];
  }
  static int out_field(int s) {
    static const int m[N_OUT] = {CLOUDSC_F_TENDENCY_LOC_T, CLOUDSC_F_TENDENCY_LOC_Q, CLOUDSC_F_TENDENCY_LOC_A,
      CLOUDSC_F_TENDENCY_LOC_CLD, CLOUDSC_F_TENDENCY_LOC_CLD, CLOUDSC_F_TENDENCY_LOC_CLD, CLOUDSC_F_TENDENCY_LOC_CLD,
      CLOUDSC_F_TENDENCY_LOC_CLD, CLOUDSC_F_PCOVPTOT, CLOUDSC_F_PLUDE,
      CLOUDSC_F_PFSQLF, CLOUDSC_F_PFSQIF, CLOUDSC_F_PFCQNNG, CLOUDSC_F_PFCQLNG, CLOUDSC_F_PFSQRF, CLOUDSC_F_PFSQSF,
      CLOUDSC_F_PFCQRNG, CLOUDSC_F_PFCQSNG, CLOUDSC_F_PFSQLTUR, CLOUDSC_F_PFSQITUR, CLOUDSC_F_PFPLSL,
      CLOUDSC_F_PFPLSN, CLOUDSC_F_PFHPSL, CLOUDSC_F_PFHPSN};
    return m[s];
  }
  template <int S> T ld(int k) const {
    const int sp = (S >= IN_PCLV0 && S <= IN_PCLV3) ? S - IN_PCLV0 : (S >= IN_TCLD0 && S <= IN_TCLD3) ? S - IN_TCLD0 : 0;
    return ((const T *)f[in_field(S)])[(size_t)(sp * klev + k) * stride + jl];
  }
  template <int S> T ld_next(int k) const { return ld<S>(k + 1); }
  template <int A> T ld_aer(int k) const {
    const int fi = A == AER_RE_ICE ? CLOUDSC_F_PRE_ICE : A == AER_ICRIT ? CLOUDSC_F_PICRIT_AER : CLOUDSC_F_PNICE;
    return ((const T *)f[fi])[(size_t)k * stride + jl];
  }
  T paph_row(int r) const { return ld<IN_PAPH>(r); }
  template <int S> void st(int k, T v) {
    const int sp = (S >= OUT_TCLD0 && S <= OUT_TCLD4) ? S - OUT_TCLD0 : 0;
    ((T *)f[out_field(S)])[(size_t)(sp * klev + k) * stride + jl] = v;
  }
  template <int S> void st_row0(T v) { st<S>(0, v); }
  T plsm() const { return ((const T *)f[CLOUDSC_F_PLSM])[jl]; }
  int ktype() const { return ((const int *)f[CLOUDSC_F_KTYPE])[jl]; }
  void st_rainfrac(T v) { ((T *)f[CLOUDSC_F_PRAINFRAC_TOPRFZ])[jl] = v; }
  void level_begin(int) {}
  void level_end(int) {}
};

/* The same policy with per-thread scratch, as the TMA kernels have it: cloudsc_column() then takes its "parked" code
 * path (flux sums, precipitation carry, per-column constants through IO::scr instead of registers). */
template <typename T>
struct HostIOScratch : HostIO<T> {
  static constexpr int kScratch = PARK_COUNT;
  T slots[PARK_COUNT];
  T &scr(int i) { return slots[i]; }
};

/* ... and with the level-range capability of the persistent kernel: the sweep of a column is cut at given levels, the
 * state at a cut (Carry registers + parked slots) goes through a buffer, and the scratch is poisoned in between -- what
 * another column tile occupying the warp's shared memory in the meantime would do to it. */
template <typename T>
struct HostIORanged : HostIOScratch<T> {
  static constexpr bool kRanged = true;
  int kb, ke;
  T *buf;   /* 8 + PARK_COUNT values */
  int k_begin() const { return kb; }
  int k_end() const { return ke; }
  void save_carry(const Carry<T> &s) {
    buf[0] = s.tp1_m1; buf[1] = s.a_m1; buf[2] = s.pap_m1; buf[3] = s.paph_k; buf[4] = s.mf_k; buf[5] = s.anewm1;
    buf[6] = s.qlnm1; buf[7] = s.qinm1;
    for (int i = 0; i < PARK_COUNT; i++) buf[8 + i] = this->slots[i];
  }
  void load_carry(Carry<T> &s) {
    s.tp1_m1 = buf[0]; s.a_m1 = buf[1]; s.pap_m1 = buf[2]; s.paph_k = buf[3]; s.mf_k = buf[4]; s.anewm1 = buf[5];
    s.qlnm1 = buf[6]; s.qinm1 = buf[7];
    for (int i = 0; i < PARK_COUNT; i++) this->slots[i] = buf[8 + i];
  }
};

template <typename T>
int run_ranged(const struct TECLDP *p, const struct cloudsc_consts *k, void *const fields[CLOUDSC_NFIELDS],
               int ngptot, int nproma, int klev, double ptsphy, int cut1, int cut2) {
  if (check_params(p) || aerosol_mask(p)) return CLOUDSC_B200_EPARAMS;
  if (!(0 < cut1 && cut1 <= cut2 && cut2 <= klev)) return CLOUDSC_B200_EARG;
  CloudscConst<T> c;
  fill_consts<T>(p, k, &c);
  StepConst<T> sc;
  fill_step<T>(c, ptsphy, &sc);
  const int nblocks = (ngptot + nproma - 1) / nproma;
  const int cuts[4] = {0, cut1, cut2, klev};
#pragma omp parallel for schedule(static)
  for (int b = 0; b < nblocks; b++) {
    void *pb[CLOUDSC_NFIELDS];
    for (int f = 0; f < CLOUDSC_NFIELDS; f++) {
      const size_t off = (size_t)b * cloudsc_b200_field_block_elems(f, nproma, klev);
      const size_t es = cloudsc_b200_field_kind(f) == 'I' ? sizeof(int) : sizeof(T);
      pb[f] = fields[f] ? (void *)((char *)fields[f] + off * es) : nullptr;
    }
    const int bsize = (ngptot - b * nproma) < nproma ? (ngptot - b * nproma) : nproma;
    for (int jl = 0; jl < bsize; jl++) {
      T buf[8 + PARK_COUNT];
      for (int r = 0; r < 3; r++) {
        if (cuts[r] == cuts[r + 1]) continue;
        HostIORanged<T> io;
        io.f = pb; io.jl = jl; io.stride = nproma; io.klev = klev;
        io.kb = cuts[r]; io.ke = cuts[r + 1]; io.buf = buf;
        for (int i = 0; i < PARK_COUNT; i++) io.slots[i] = T(-1.0e300);
        cloudsc_column<T>(c, sc, io, klev);
      }
    }
  }
  return 0;
}

/* The split-tail launch (schedule mode 3) on a pretend machine of `nsm` SMs with `per_sm` resident one-warp CTAs: the blocks of
 * the launch are walked in index order through split_piece() / split_count() / split_cut() -- the plan the device runs
 * (cloudsc_split_plan.h) -- each block sweeping the 32 columns of its tile over its level range.  The parked state of a split
 * tile lives in `carry[tile]`, its flag is set by the upper piece and REQUIRED by the lower one (a plan that ran a lower
 * piece first would return -2 here and spin on the device); scratch is poisoned at every block start, and so is a carry
 * once it was consumed.  info[4] = { blocks, nsplit, cut, pieces that loaded a carry }. */
template <typename T>
int run_split(const struct TECLDP *p, const struct cloudsc_consts *k, void *const fields[CLOUDSC_NFIELDS],
              int ngptot, int nproma, int klev, double ptsphy, int nsm, int per_sm, int ksplit, int target_b, int info[4]) {
  if (check_params(p) || aerosol_mask(p)) return CLOUDSC_B200_EPARAMS;
  if (nsm < 1 || per_sm < 1 || klev < 3) return CLOUDSC_B200_EARG;
  CloudscConst<T> c;
  fill_consts<T>(p, k, &c);
  StepConst<T> sc;
  fill_step<T>(c, ptsphy, &sc);
  const int ntiles = (ngptot + 31) / 32;
  const int nsplit = split_count(ntiles, nsm, per_sm, target_b, 4096, ntiles);
  const int cut = split_cut(ksplit, klev);
  const int nblocks = split_blocks(ntiles, nsplit);
  constexpr int kVals = 8 + PARK_COUNT;
  std::vector<T> carry((size_t)nsplit * 32 * kVals, T(-7.0e300));
  std::vector<int> flag(nsplit, 0);
  std::vector<int> covered((size_t)ntiles * klev, 0);
  int loads = 0;
  for (int b = 0; b < nblocks; b++) {
    const SplitPiece pc = split_piece(b, ntiles, nsplit, cut, klev);
    if (pc.tile < 0 || pc.tile >= ntiles || pc.kb < 0 || pc.ke > klev || pc.kb >= pc.ke) return -3;
    if (pc.lower) {
      if (pc.tile >= nsplit || flag[pc.tile] != 1) return -2;   /* the upper piece has not run */
      flag[pc.tile] = 2;
      ++loads;
    }
    for (int kk = pc.kb; kk < pc.ke; kk++) ++covered[(size_t)pc.tile * klev + kk];
    for (int lane = 0; lane < 32; lane++) {
      const int col = pc.tile * 32 + lane;
      if (col >= ngptot) break;
      const int blk = col / nproma, jl = col - blk * nproma;
      void *pb[CLOUDSC_NFIELDS];
      for (int f = 0; f < CLOUDSC_NFIELDS; f++) {
        const size_t off = (size_t)blk * cloudsc_b200_field_block_elems(f, nproma, klev);
        const size_t es = cloudsc_b200_field_kind(f) == 'I' ? sizeof(int) : sizeof(T);
        pb[f] = fields[f] ? (void *)((char *)fields[f] + off * es) : nullptr;
      }
      HostIORanged<T> io;
      io.f = pb; io.jl = jl; io.stride = nproma; io.klev = klev;
      io.kb = pc.kb; io.ke = pc.ke;
      io.buf = (pc.upper || pc.lower) ? &carry[((size_t)pc.tile * 32 + lane) * kVals] : nullptr;
      for (int i = 0; i < PARK_COUNT; i++) io.slots[i] = T(-1.0e300);
      cloudsc_column<T>(c, sc, io, klev);
      if (pc.lower) for (int i = 0; i < kVals; i++) io.buf[i] = T(-7.0e300);
    }
    if (pc.upper) {
      if (flag[pc.tile] != 0) return -4;
      flag[pc.tile] = 1;
    }
  }
  for (int t = 0; t < nsplit; t++) if (flag[t] != 2) return -5;                  /* every parked state was consumed */
  for (size_t i = 0; i < covered.size(); i++) if (covered[i] != 1) return -6;   /* every level of every tile exactly once */
  if (info) { info[0] = nblocks; info[1] = nsplit; info[2] = cut; info[3] = loads; }
  return 0;
}

template <typename T, bool PARKED = false>
int run(const struct TECLDP *p, const struct cloudsc_consts *k, void *const fields[CLOUDSC_NFIELDS],
        int ngptot, int nproma, int klev, double ptsphy) {
  if (check_params(p)) return CLOUDSC_B200_EPARAMS;
  CloudscConst<T> c;
  fill_consts<T>(p, k, &c);
  StepConst<T> sc;
  fill_step<T>(c, ptsphy, &sc);
  const int nblocks = (ngptot + nproma - 1) / nproma;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < nblocks; b++) {
    void *pb[CLOUDSC_NFIELDS];
    for (int f = 0; f < CLOUDSC_NFIELDS; f++) {
      const size_t off = (size_t)b * cloudsc_b200_field_block_elems(f, nproma, klev);
      const size_t es = cloudsc_b200_field_kind(f) == 'I' ? sizeof(int) : sizeof(T);
      pb[f] = fields[f] ? (void *)((char *)fields[f] + off * es) : nullptr;
    }
    const int bsize = (ngptot - b * nproma) < nproma ? (ngptot - b * nproma) : nproma;
    for (int jl = 0; jl < bsize; jl++) {
      if (PARKED) {
        HostIOScratch<T> io;
        io.f = pb; io.jl = jl; io.stride = nproma; io.klev = klev;
        for (int i = 0; i < PARK_COUNT; i++) io.slots[i] = T(-1.0e300);   /* poison: every slot must be written first */
        cloudsc_column<T>(c, sc, io, klev);
      } else if (aerosol_mask(p)) {   /* the instantiation the device launches when laericesed / laericeauto are set */
        HostIO<T> io{pb, jl, nproma, klev};
        cloudsc_column<T, true>(c, sc, io, klev);
      } else {
        HostIO<T> io{pb, jl, nproma, klev};
        cloudsc_column<T>(c, sc, io, klev);
      }
    }
  }
  return 0;
}
}  // namespace

extern "C" {
/* local copies of the two table helpers so this test library has no dependency on the product .so */
static const char k_kinds[] = {
#define CLOUDSC_X(name, kind, use) #kind[0],
  CLOUDSC_FIELDS(CLOUDSC_X)
#undef CLOUDSC_X
  0};
char cloudsc_b200_field_kind(int f) { return k_kinds[f]; }
size_t cloudsc_b200_field_block_elems(int f, int nproma, int klev) {
  switch (k_kinds[f]) {
    case 'L': return (size_t)klev * nproma;
    case 'H': return (size_t)(klev + 1) * nproma;
    case 'C': return (size_t)5 * klev * nproma;
    default: return (size_t)nproma;
  }
}
int cloudsc_emul_blocks_f64(const struct TECLDP *p, const struct cloudsc_consts *k, void *const fields[CLOUDSC_NFIELDS],
                            int ngptot, int nproma, int klev, double ptsphy) {
  return run<double>(p, k, fields, ngptot, nproma, klev, ptsphy);
}
int cloudsc_emul_blocks_parked_f64(const struct TECLDP *p, const struct cloudsc_consts *k, void *const fields[CLOUDSC_NFIELDS],
                                   int ngptot, int nproma, int klev, double ptsphy) {
  return run<double, true>(p, k, fields, ngptot, nproma, klev, ptsphy);
}
int cloudsc_emul_blocks_ranged_f64(const struct TECLDP *p, const struct cloudsc_consts *k, void *const fields[CLOUDSC_NFIELDS],
                                   int ngptot, int nproma, int klev, double ptsphy, int cut1, int cut2) {
  return run_ranged<double>(p, k, fields, ngptot, nproma, klev, ptsphy, cut1, cut2);
}
int cloudsc_emul_split_tail_f64(const struct TECLDP *p, const struct cloudsc_consts *k, void *const fields[CLOUDSC_NFIELDS],
                                int ngptot, int nproma, int klev, double ptsphy, int nsm, int per_sm, int ksplit, int target_b,
                                int info[4]) {
  return run_split<double>(p, k, fields, ngptot, nproma, klev, ptsphy, nsm, per_sm, ksplit, target_b, info);
}
/* the plan alone: out[5] = { tile, kb, ke, lower, upper } of block b */
void cloudsc_emul_split_piece(int b, int ntiles, int nsplit, int cut, int klev, int out[5]) {
  const SplitPiece pc = split_piece(b, ntiles, nsplit, cut, klev);
  out[0] = pc.tile; out[1] = pc.kb; out[2] = pc.ke; out[3] = pc.lower; out[4] = pc.upper;
}
int cloudsc_emul_split_count(int ntiles, int nsm, int per_sm, int target_b, int max_flags, int carry_tiles) {
  return split_count(ntiles, nsm, per_sm, target_b, max_flags, carry_tiles);
}
int cloudsc_emul_split_cut(int ksplit, int klev) { return split_cut(ksplit, klev); }
/* ---- per-process entry points (tests/test_unit_processes.py): the device code's sink limiter, implicit solve and
 * Newton step, called on their own.  a[13] = a01 a02 a03 a04 a12 a14 a23 a24 a34 d0 d1 d2 d3 (cloudsc_column.cuh). */
void cloudsc_emul_unit_limiter(double a[13], const double qx5[5], double cs[5]) {
  CloudscConst<double> c;
  memset(&c, 0, sizeof(c));
  c.zepsec = 1.0e-14;
  const double qx[4] = {qx5[0], qx5[1], qx5[2], qx5[3]};
  sink_limiter<double>(c, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], qx, qx5[4],
                       cs[0], cs[1], cs[2], cs[3], cs[4]);
}
void cloudsc_emul_unit_solve(const double qx5[5], const double cs[5], const double qb[4], const double fallsink[3], double qxn_out[5]) {
  CloudscConst<double> c;
  memset(&c, 0, sizeof(c));
  c.zepsec = 1.0e-14;
  const double qx[4] = {qx5[0], qx5[1], qx5[2], qx5[3]};
  double qxn[5];
  implicit_solve<double>(c, qx, qx5[4], cs[0], cs[1], cs[2], cs[3], cs[4], qb[0], qb[1], qb[2], qb[3], fallsink[0], fallsink[1],
                         fallsink[2], qxn);
  for (int i = 0; i < 5; i++) qxn_out[i] = qxn[i];
}
int cloudsc_emul_unit_satadj(const struct TECLDP *p, const struct cloudsc_consts *k, double qp, double *tn, double *qsm) {
  if (check_params(p)) return CLOUDSC_B200_EPARAMS;
  CloudscConst<double> c;
  fill_consts<double>(p, k, &c);
  exp2_tab_init();
  satadj_step<double>(c, qp, *tn, *qsm);
  return 0;
}
int cloudsc_emul_blocks_f32(const struct TECLDP *p, const struct cloudsc_consts *k, void *const fields[CLOUDSC_NFIELDS],
                            int ngptot, int nproma, int klev, double ptsphy) {
  return run<float>(p, k, fields, ngptot, nproma, klev, ptsphy);
}
}
