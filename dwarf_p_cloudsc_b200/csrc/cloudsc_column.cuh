/*
 * cloudsc_column.cuh -- CLOUDSC for ONE column as a single fused top-to-bottom sweep.
 *
 * This is the arithmetic of the reference kernel src/cloudsc_c/cloudsc/cloudsc_c.c:360-2582
 * (dwarf-cloudsc-c) re-organised for one GPU thread per column:
 *
 *  - ONE loop over the klev levels.  The reference makes ~10 passes over [klev][klon] temporaries
 *    (init :462-513, tidy-up :519-575, saturation :583-640, sweep :732-2508, fluxes :2521-2582,
 *    ~51 KB of stack per column); here every such temporary is a scalar of the current level and
 *    the only state that survives from level k-1 to k is `struct Carry` below.
 *  - the 5x5 source/sink matrix is a fully unrolled register array; entries no process ever writes
 *    stay compile-time zeros.  With the shipped process set (iwarmrain=2) ZSOLQB has only the
 *    entries (l,l) (i,i) (l,s) (i,s), so the non-pivoting LU of :2351-2397 collapses to two
 *    multipliers and one divide per species (derivation at "5.2.2" below; value-identical to the
 *    full elimination, signs of zeros aside).  The sink limiter (:2233-2286) only divides when a
 *    species' sinks really exceed its content (ratio == 1 otherwise, exactly as in the reference).
 *  - every output element is written exactly once (the reference read-modify-writes the
 *    tendencies, :526-567 then :2465-2498, and re-reads the flux sums, :2548-2568).
 *  - parameters come from `CloudscConst<T>` (__constant__ memory) and `StepConst<T>` (kernel
 *    parameter bank); no parameter is ever read from global memory.
 *
 * Two arithmetic modes, selected at compile time:
 *
 *   CLOUDSC_STRICT=1  statement-by-statement evaluation order of the reference (comments give the
 *                     cloudsc_c.c line).  A host build of this header in this mode with FP
 *                     contraction off is BIT-IDENTICAL to the reference on the same inputs
 *                     (tests/host_emul, tests/test_cpu_oracle.py) -- that is the proof that the
 *                     k-caching / sparse-solve re-organisation preserves the algorithm.
 *   CLOUDSC_STRICT=0  (the product kernels) same algorithm with shared reciprocals
 *                     (a/b and c/b -> r=1/b; a*r; c*r), a Newton reciprocal without the IEEE
 *                     special-case path and a branch-free exp() -- each within ~2 ulp of the strict
 *                     operation.  tests/ bound the end-to-end effect (CPU: fast-vs-reference;
 *                     GPU: the dwarf's error table within BASELINE.md's tolerance).
 *
 * Generic over the arithmetic type T (double, float) and over an IO policy providing
 * level-indexed loads/stores.  Physics switches are specialised on the values the reference
 * hard-codes (cloudsc_c.c:367-382: iwarmrain=2, ievaprain=2, ievapsnow=1, idepice=1) and on
 * laer*=false; cloudsc_b200_set_params() refuses any other parameter set.
 */
#ifndef CLOUDSC_COLUMN_CUH
#define CLOUDSC_COLUMN_CUH

#include "cloudsc_consts.h"

#ifndef CLOUDSC_STRICT
#define CLOUDSC_STRICT 0
#endif

#if defined(__CUDACC__)
#define CLOUDSC_FN __device__ __forceinline__
#else
#include <cmath>
#include <cstring>
#define CLOUDSC_FN static inline
#endif

/* Code-layout hint only: blocks a minority of grid points reach (warm-phase and precipitation
 * processes, detrainment, prescribed supersaturation) are moved out of the fall-through path, so the
 * common path is fetched sequentially (the level body is ~3x the 32 KB instruction cache). */
#if defined(__GNUC__) || defined(__CUDACC__)
#define CLOUDSC_RARE(x) __builtin_expect(!!(x), 0)
#else
#define CLOUDSC_RARE(x) (x)
#endif

namespace cloudsc {

/* ---- input / output slots the IO policy understands ----------------------------------------- */
enum InSlot {
  IN_PT, IN_PQ, IN_TT, IN_TQ, IN_TA, IN_PA, IN_PAP, IN_PSUPSAT, IN_PLUDE, IN_PLU, IN_PSNDE,
  IN_PMFU, IN_PMFD, IN_PHRSW, IN_PHRLW, IN_PVERVEL, IN_PVFL, IN_PVFI,
  IN_PCLV0, IN_PCLV1, IN_PCLV2, IN_PCLV3, IN_TCLD0, IN_TCLD1, IN_TCLD2, IN_TCLD3,
  IN_PAPH, /* [klev+1] */
  N_IN
};
enum OutSlot {
  OUT_TT, OUT_TQ, OUT_TA, OUT_TCLD0, OUT_TCLD1, OUT_TCLD2, OUT_TCLD3, OUT_TCLD4, OUT_COVPTOT,
  OUT_PLUDE,
  /* [klev+1] flux rows */
  OUT_FSQLF, OUT_FSQIF, OUT_FCQNNG, OUT_FCQLNG, OUT_FSQRF, OUT_FSQSF, OUT_FCQRNG, OUT_FCQSNG,
  OUT_FSQLTUR, OUT_FSQITUR, OUT_FPLSL, OUT_FPLSN, OUT_FHPSL, OUT_FHPSN,
  N_OUT
};

/* aerosol inputs, read only by the AER instantiation of cloudsc_column() and only where the matching flag is set */
enum AerSlot { AER_RE_ICE, AER_ICRIT, AER_NICE };

/* ---- math wrappers: same names for double/float, host/device -------------------------------- */
/* max/min as compare+select.  For the NaN-free values of this scheme they return what
 * fmax/fmin return (zeros of either sign compare equal), at 3 instead of 7 instructions on
 * sm_100a (there is no DMNMX; fmax() expands to a NaN-aware sequence). */
template <typename T> CLOUDSC_FN T mx(T a, T b) { return a > b ? a : b; }
template <typename T> CLOUDSC_FN T mn(T a, T b) { return a < b ? a : b; }
#if defined(__CUDACC__)
/* single precision has a one-instruction max/min (FMNMX) */
CLOUDSC_FN float mx(float a, float b) { return fmaxf(a, b); }
CLOUDSC_FN float mn(float a, float b) { return fminf(a, b); }
#endif
#if defined(__CUDACC__)
CLOUDSC_FN double ex_lib(double x) { return exp(x); }
CLOUDSC_FN float ex_lib(float x) { return expf(x); }
CLOUDSC_FN double pw(double x, double y) { return pow(x, y); }
CLOUDSC_FN float pw(float x, float y) { return powf(x, y); }
CLOUDSC_FN double sq(double x) { return sqrt(x); }
CLOUDSC_FN float sq(float x) { return sqrtf(x); }
CLOUDSC_FN double ab(double x) { return fabs(x); }
CLOUDSC_FN float ab(float x) { return fabsf(x); }
CLOUDSC_FN double cs(double x, double y) { return copysign(x, y); }
CLOUDSC_FN float cs(float x, float y) { return copysignf(x, y); }
#else
static inline double ex_lib(double x) { return std::exp(x); }
static inline float ex_lib(float x) { return std::exp(x); }
static inline double pw(double x, double y) { return std::pow(x, y); }
static inline float pw(float x, float y) { return std::pow(x, y); }
static inline double sq(double x) { return std::sqrt(x); }
static inline float sq(float x) { return std::sqrt(x); }
static inline double ab(double x) { return std::fabs(x); }
static inline float ab(float x) { return std::fabs(x); }
static inline double cs(double x, double y) { return std::copysign(x, y); }
static inline float cs(float x, float y) { return std::copysign(x, y); }
#endif

/* ---- fast-mode primitives -------------------------------------------------------------------- */
/*
 * exp(x) without the library's special-case handling, for |x| <= 700: n = rint(x*log2 e),
 * r = x - n ln2 (two-constant Cody-Waite), degree-12 polynomial on |r| <= ln2/2 (fitted in extended
 * precision, max rel. error 1.5e-16), scaling by adding n to the exponent field.  ~24 instructions
 * instead of ~66.  Call sites whose argument is bounded by the physics (saturation formulae with
 * 160 K <= T, nucleation, linear functions of T - RTT) use it directly; exp_fast_lo() clamps from
 * below for the two that can run away to -inf (exp(-x^2) of the snow autoconversion, and pow_fast of
 * a vanishing base), where the result underflows harmlessly to 1e-304 instead of 0.
 */
#ifndef CLOUDSC_EXP_TABLE
#define CLOUDSC_EXP_TABLE 1
#endif
#ifndef CLOUDSC_EXP_ESTRIN
#define CLOUDSC_EXP_ESTRIN 0
#endif
#if CLOUDSC_EXP_TABLE
/* Table-driven variant (the default): x = (64 n + j) ln2/64 + r, exp x = 2^n 2^(j/64) (1 + expm1 r) with a 64-entry
 * table of correctly rounded 2^(j/64) and the degree-5 Taylor polynomial on |r| <= ln2/128 (truncation 3.5e-17,
 * total ~1 ulp).  15 instructions and 8 constants instead of 25 and 16: instruction supply is one of the limits of
 * the FP64 kernel, so an exp costs its instructions and its code bytes.  On the device the table is
 * staged in shared memory by every warp before the sweep (exp2_tab_init); a lane reads its entry with one LDS. */
#define CLOUDSC_EXP2_TAB_VALUES \
  0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0, \
  0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0, \
  0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0, \
  0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0, \
  0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0, \
  0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0, \
  0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0, \
  0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0, \
  0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0, \
  0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0, \
  0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0, \
  0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0, \
  0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0, \
  0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0, \
  0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0, \
  0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0 \

#if defined(__CUDACC__)
static __constant__ double k_exp2_tab_c[64] = {CLOUDSC_EXP2_TAB_VALUES};
__device__ __forceinline__ double *exp2_tab() {
  __shared__ double tab[64];
  return tab;
}
/* every warp writes all 64 entries (identical values: no ordering between warps is needed); the entries are dealt
 * out over the lanes that are actually here, so a partial warp (ragged last block) fills the whole table too */
__device__ __forceinline__ void exp2_tab_init() {
  const unsigned m = __activemask();
  const int lane = (int)(threadIdx.x & 31u);
  const int n = __popc(m), rank = __popc(m & ((1u << lane) - 1u));
  double *tab = exp2_tab();
  for (int i = rank; i < 64; i += n) tab[i] = k_exp2_tab_c[i];
  __syncwarp(m);
}
#else
static const double k_exp2_tab_h[64] = {CLOUDSC_EXP2_TAB_VALUES};
static inline void exp2_tab_init() {}
#endif
CLOUDSC_FN double exp_fast(double x) {
  const double magic = 6755399441055744.0; /* 1.5 * 2^52 */
  const double t = x * 92.332482616893658 + magic;             /* 64 / ln 2 */
  const double fn = t - magic;
  double r = fn * (-6.93147180369123816490e-01 / 64.0) + x;    /* ln2_hi / 64: 21 significant bits, fn < 2^17 */
  r = fn * (-1.90821492927058770002e-10 / 64.0) + r;
#if CLOUDSC_EXP_ESTRIN
  /* (c5 r + c4) r^2 + (c3 r + c2): the same three FMAs, two deep instead of three */
  const double r2 = r * r;
  const double q = (r * 8.3333333333333332e-03 + 4.1666666666666664e-02) * r2 + (r * 1.6666666666666666e-01 + 0.5);
  const double p = r2 * q + r;                                 /* expm1(r) */
#else
  double q = r * 8.3333333333333332e-03 + 4.1666666666666664e-02;
  q = r * q + 1.6666666666666666e-01;
  q = r * q + 0.5;
  const double p = (r * r) * q + r;                            /* expm1(r) */
#endif
#if defined(__CUDACC__)
  const int ti = __double2loint(t);
  const double tv = exp2_tab()[ti & 63];
  const double res = tv * p + tv;
  return __hiloint2double(__double2hiint(res) + ((ti >> 6) << 20), __double2loint(res));
#else
  long long tb, pb;
  std::memcpy(&tb, &t, 8);
  const int ti = (int)(tb & 0xffffffffLL);
  const double tv = k_exp2_tab_h[ti & 63];
  double res = tv * p + tv;
  std::memcpy(&pb, &res, 8);
  pb += (long long)(ti >> 6) << 52;
  std::memcpy(&res, &pb, 8);
  return res;
#endif
}
#else
CLOUDSC_FN void exp2_tab_init() {}
CLOUDSC_FN double exp_fast(double x) {
  const double magic = 6755399441055744.0; /* 1.5 * 2^52 */
  const double t = x * 1.4426950408889634 + magic;
  const double fn = t - magic;
  double r = fn * -6.93147180369123816490e-01 + x;
  r = fn * -1.90821492927058770002e-10 + r;
  /* Estrin evaluation: 4 dependent FMAs deep instead of 13 (DFMA latency is ~8 cycles on sm_100a and
     only 3 warps per scheduler are resident) */
  const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
  const double p01 = r + 1.0;
  const double p23 = 0.1666666666666667 * r + 0.5;
  const double p45 = 0.00833333333332612 * r + 0.04166666666666668;
  const double p67 = 0.0001984126987497067 * r + 0.0013888888888879635;
  const double p89 = 2.7557255115742253e-06 * r + 2.480158733369391e-05;
  const double pab = 2.5105361442549964e-08 * r + 2.7557266270270763e-07;
  const double q0 = p23 * r2 + p01;
  const double q1 = p67 * r2 + p45;
  const double q2 = pab * r2 + p89;
  const double s0 = q1 * r4 + q0;
  const double s1 = 2.09172269697194e-09 * r4 + q2;
  double p = s1 * r8 + s0;
#if defined(__CUDACC__)
  const int n = __double2loint(t);
  return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
#else
  long long tb, pb;
  std::memcpy(&tb, &t, 8);
  std::memcpy(&pb, &p, 8);
  pb += (long long)(int)(tb & 0xffffffffLL) << 52;
  std::memcpy(&p, &pb, 8);
  return p;
#endif
}
#endif
/* Single precision: the special-function unit is accurate to ~2 ulp, which is all a 24-bit result can use.
 * expf() is already branch-free (range reduction + MUFU.EX2); division, sqrt and powf are the library's
 * branchy IEEE sequences (64 slow-path calls in the FP32 kernel), replaced below by MUFU.RCP + one Newton
 * step, MUFU.SQRT and ex2(y lg2 x). */
CLOUDSC_FN float exp_fast(float x) { return ex_lib(x); }
/* max(x, -700): -700.0 is 0xC085E000'00000000, so x < -700 (or == -700) exactly when the high word of x, read as an
 * unsigned integer, is >= 0xC085E000 -- one ISETP instead of a DSETP in front of every pow */
#ifndef CLOUDSC_EXPLO_INT
#define CLOUDSC_EXPLO_INT 1
#endif
CLOUDSC_FN double exp_fast_lo(double x) {
#if defined(__CUDACC__) && CLOUDSC_EXPLO_INT
  return exp_fast((unsigned)__double2hiint(x) >= 0xC085E000u ? -700.0 : x);
#else
  return exp_fast(mx(x, -700.0));
#endif
}
CLOUDSC_FN float exp_fast_lo(float x) { return ex_lib(x); }

/* 1/x by the hardware seed (MUFU.RCP64H, ~2^-20) + one cubically convergent step
 * r(1 + e + e^2), e = 1 - x r  (error e^3 ~ 2^-60), no zero/inf/denormal path (every divisor of the
 * scheme is a normal number bounded away from 0 by an explicit max(., eps) or by its physics). */
CLOUDSC_FN double rcp_fast(double x) {
#if defined(__CUDACC__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  const double p = fma(e, e, e);
  return fma(r, p, r);
#else
  return 1.0 / x;
#endif
}
CLOUDSC_FN float rcp_fast(float x) {
#if defined(__CUDACC__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(r, fmaf(-x, r, 1.0f), r);
#else
  return 1.0f / x;
#endif
}
/* a / b where the reciprocal itself is not needed again: the quotient a r0 is refined directly, q = q0 + q0 (e + e^2) with
 * q0 = a r0 evaluated NEXT TO e = 1 - b r0 -- three dependent FP64 operations behind the MUFU seed instead of four */
#ifndef CLOUDSC_DIV_FUSED
#define CLOUDSC_DIV_FUSED 1
#endif
CLOUDSC_FN double div_fast(double a, double b) {
#if defined(__CUDACC__) && CLOUDSC_DIV_FUSED
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  const double e = fma(-b, r, 1.0);
  const double q0 = a * r;
  const double p = fma(e, e, e);
  return fma(q0, p, q0);
#else
  return a * rcp_fast(b);
#endif
}
CLOUDSC_FN float div_fast(float a, float b) { return a * rcp_fast(b); }

/*
 * log(x) for normal positive x: x = m 2^e with m in [sqrt(1/2), sqrt(2)), s = (m-1)/(m+1),
 * log m = 2 atanh s = 2s + 2 s z g(z), z = s^2, g a degree-6 polynomial (tools/fit_log.py, fit error
 * 4e-18 relative to log m).  ~35 instructions.  x = 0 / denormal return about -709 (callers only use it
 * under exp_fast(y * log x), where that underflows to ~0 like the exact result).
 */
CLOUDSC_FN double log_fast(double x) {
#if defined(__CUDACC__)
  int hi = __double2hiint(x);
  const int lo = __double2loint(x);
#else
  long long xb;
  std::memcpy(&xb, &x, 8);
  int hi = (int)(xb >> 32);
  const int lo = (int)(xb & 0xffffffffLL);
#endif
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  if (hi >= 0x3ff6a09f) { hi -= 0x00100000; e += 1; }
#if defined(__CUDACC__)
  const double m = __hiloint2double(hi, lo);
#else
  double m;
  xb = ((long long)hi << 32) | (long long)(unsigned)lo;
  std::memcpy(&m, &xb, 8);
#endif
  const double s = div_fast(m - 1.0, m + 1.0);
  const double z = s * s;
  const double z2 = z * z;
  const double g01 = 0.19999999999949722 * z + 0.3333333333333335;
  const double g23 = 0.11111105565129939 * z + 0.14285714313001635;
  const double g45 = 0.07665855309105386 * z + 0.09091444727837474;
  const double g46 = 0.07308292712444311 * z2 + g45;
  const double fe = (double)e;
  const double g = (g46 * z2 + g23) * z2 + g01;
  const double lm = (s * z) * g + s;           /* atanh s */
  return fe * 6.93147180369123816490e-01 + (fe * 1.90821492927058770002e-10 + (lm + lm));
}
/* x^y for x > 0 as exp(y log x): relative error ~ |y log x| * 2e-16 (library pow: 0.5 ulp at ~15x the
 * instructions and a call into a spilling slow path).  exp_fast is declared above. */
CLOUDSC_FN double pow_fast(double x, double y) { return exp_fast_lo(y * log_fast(x)); }
/* x^y = 2^(y lg2 x) on the special-function unit: relative error ~ |y lg2 x| * 2^-22; x = 0 gives 2^-inf = 0 */
CLOUDSC_FN float pow_fast(float x, float y) {
#if defined(__CUDACC__)
  float l, r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y * l));
  return r;
#else
  return pw(x, y);
#endif
}
/* sqrt(x) for normal x > 0: MUFU.RSQ64H seed, two Newton steps on 1/sqrt, one on the root */
#ifndef CLOUDSC_SQRT_CUBIC
#define CLOUDSC_SQRT_CUBIC 1
#endif
CLOUDSC_FN double sqrt_fast(double x) {
#if defined(__CUDACC__) && CLOUDSC_SQRT_CUBIC
  /* one cubically convergent step on the root itself: t = x y0 ~ sqrt x, eps = 1 - x y0^2 (~2^-20 for the MUFU.RSQ64H seed),
     sqrt x = t (1 - eps)^(-1/2) = t (1 + eps/2 + 3 eps^2 / 8 + O(eps^3 ~ 2^-60)): four dependent operations instead of nine */
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double t = x * y;
  const double eps = fma(-t, y, 1.0);
  const double p = fma(eps, 0.375, 0.5);
  return fma(t * eps, p, t);
#elif defined(__CUDACC__)
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double h = 0.5 * x;
  double e = fma(-h * r, r, 0.5);
  r = fma(r, e, r);
  e = fma(-h * r, r, 0.5);
  r = fma(r, e, r);
  const double q = x * r;
  return fma(fma(-q, q, x), 0.5 * r, q);
#else
  return std::sqrt(x);
#endif
}
CLOUDSC_FN float sqrt_fast(float x) {
#if defined(__CUDACC__)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sq(x);
#endif
}

/* max(x, 0).  Fast mode, double, device: a test of the sign bit on the integer pipe (ISETP + 2 SEL) instead of a compare
 * on the FP64 pipe (DSETP + 2 FSEL); same value for every non-NaN x (-0 gives +0 either way). */
#ifndef CLOUDSC_MX0_INT
#define CLOUDSC_MX0_INT 1
#endif
CLOUDSC_FN double mx0_fast(double x) {
#if defined(__CUDACC__) && CLOUDSC_MX0_INT
  return __double2hiint(x) < 0 ? 0.0 : x;
#else
  return x > 0.0 ? x : 0.0;
#endif
}
CLOUDSC_FN float mx0_fast(float x) { return mx(x, 0.0f); }
/* min(x, C) / max(x, C) for a constant C > 0 whose low word is zero (1.0, 0.5, 160.0): x < C exactly when the high word of x
 * is below that of C as a signed integer (negative x included), so the compare moves to the integer pipe as well. */
template <int HI> CLOUDSC_FN double mn_hi(double x, double cval) {
#if defined(__CUDACC__) && CLOUDSC_MX0_INT
  return __double2hiint(x) < HI ? x : cval;
#else
  return x < cval ? x : cval;
#endif
}
template <int HI> CLOUDSC_FN float mn_hi(float x, float cval) { return mn(x, cval); }
template <int HI> CLOUDSC_FN double mx_hi(double x, double cval) {
#if defined(__CUDACC__) && CLOUDSC_MX0_INT
  return __double2hiint(x) < HI ? cval : x;
#else
  return x > cval ? x : cval;
#endif
}
template <int HI> CLOUDSC_FN float mx_hi(float x, float cval) { return mx(x, cval); }
#if CLOUDSC_STRICT
#define CLOUDSC_MX0(x) mx((x), zero)
#define CLOUDSC_MX0R(x) mx(zero, (x))
#define CLOUDSC_MN1(x) mn((x), T(1.0))
#define CLOUDSC_MN1R(x) mn(T(1.0), (x))
#define CLOUDSC_MNH(x) mn((x), T(0.5))
#define CLOUDSC_MNHR(x) mn(T(0.5), (x))
#define CLOUDSC_MX160(x) mx((x), T(160.0))
#else
#define CLOUDSC_MX0(x) mx0_fast(x)
#define CLOUDSC_MX0R(x) mx0_fast(x)
#define CLOUDSC_MN1(x) mn_hi<0x3ff00000>((x), T(1.0))
#define CLOUDSC_MN1R(x) mn_hi<0x3ff00000>((x), T(1.0))
#define CLOUDSC_MNH(x) mn_hi<0x3fe00000>((x), T(0.5))
#define CLOUDSC_MNHR(x) mn_hi<0x3fe00000>((x), T(0.5))
#define CLOUDSC_MX160(x) mx_hi<0x40640000>((x), T(160.0))
#endif
#if CLOUDSC_STRICT
#define CLOUDSC_RCP(x) (one / (x))
#define CLOUDSC_DIV(a, b) ((a) / (b))
#define CLOUDSC_DIVR(a, b, r) ((a) / (b))  /* a/b where r == 1/b is available in fast mode */
#define CLOUDSC_POW(x, y) pw(x, y)
#define CLOUDSC_EXP_COLD(x) ex_lib(x)
#define CLOUDSC_EXP_LO_COLD(x) ex_lib(x)
#define CLOUDSC_SQRT(x) sq(x)
template <typename T> CLOUDSC_FN T pow15(T x) { return pw(x, T(1.5)); }
template <typename T> CLOUDSC_FN T cube(T x) { return pw(x, T(3.0)); }
#define CLOUDSC_POW15(x) pow15<T>(x)
#define CLOUDSC_CUBE(x) cube<T>(x)
template <typename T> CLOUDSC_FN T ex(T x) { return ex_lib(x); }
template <typename T> CLOUDSC_FN T ex_lo(T x) { return ex_lib(x); }
#else
#define CLOUDSC_RCP(x) rcp_fast(x)
#define CLOUDSC_DIV(a, b) div_fast((a), (b))
#define CLOUDSC_DIVR(a, b, r) ((a) * (r))
/* CLOUDSC_CALLS: the FP64 pow (log + exp, ~62 instructions, 13 sites) and the exp of the process blocks as ONE
 * out-of-line copy each.  The loop body of the FP64 kernel (2400 hot instructions = 38 KB) does not fit the SM's
 * 32 KB instruction cache (ncu: sm__icc_request_hit_rate 83 %, GPC instruction-cache requests at 95 % of peak);
 * the saturation / Newton exps on the critical path stay inline. */
#ifndef CLOUDSC_CALLS
#define CLOUDSC_CALLS 0
#endif
#if CLOUDSC_CALLS && defined(__CUDACC__)
__device__ __noinline__ double pow_call(double x, double y) { return pow_fast(x, y); }
__device__ __noinline__ double exp_call(double x) { return exp_fast(x); }
__device__ __noinline__ double exp_lo_call(double x) { return exp_fast_lo(x); }
CLOUDSC_FN float pow_call(float x, float y) { return pow_fast(x, y); }
CLOUDSC_FN float exp_call(float x) { return exp_fast(x); }
CLOUDSC_FN float exp_lo_call(float x) { return exp_fast_lo(x); }
#define CLOUDSC_POW(x, y) pow_call(x, y)
#define CLOUDSC_EXP_COLD(x) exp_call(x)
#define CLOUDSC_EXP_LO_COLD(x) exp_lo_call(x)
#else
#define CLOUDSC_POW(x, y) pow_fast(x, y)
#define CLOUDSC_EXP_COLD(x) exp_fast(x)
#define CLOUDSC_EXP_LO_COLD(x) exp_fast_lo(x)
#endif
#define CLOUDSC_SQRT(x) sqrt_fast(x)
template <typename T> CLOUDSC_FN T pow15(T x) { return x * sqrt_fast(x); }
template <typename T> CLOUDSC_FN T cube(T x) { return (x * x) * x; }
#define CLOUDSC_POW15(x) pow15<T>(x)
#define CLOUDSC_CUBE(x) cube<T>(x)
template <typename T> CLOUDSC_FN T ex(T x) { return exp_fast(x); }
template <typename T> CLOUDSC_FN T ex_lo(T x) { return exp_fast_lo(x); }
#endif

/* Optional per-thread scratch of an IO policy (IO::kScratch slots, `T &scr(int)`): values that live through the whole
 * sweep but are touched once per level (the six running flux sums) or only in some blocks (per-column constants) are
 * parked there instead of holding 20 of the 168 registers.  IO policies without scratch keep them in registers. */
template <typename IO, typename = void> struct io_scratch { static constexpr int value = 0; };
template <typename IO> struct io_scratch<IO, decltype((void)IO::kScratch)> { static constexpr int value = IO::kScratch; };
enum { PARK_FSQLF, PARK_FSQIF, PARK_FCQLNG, PARK_FCQNNG, PARK_FSQLTUR, PARK_FSQITUR, PARK_PAPH_SURF, PARK_R_PAPH_SURF,
       PARK_PF_I, PARK_PF_R, PARK_PF_S, PARK_COVPTOT, PARK_COVPMAX, PARK_CLDTOPDIST, PARK_RAINFRAC, PARK_COUNT };
template <int I, typename T, typename IO> CLOUDSC_FN T park_ld(IO &io, const T &reg) {
  if constexpr (io_scratch<IO>::value > I) return io.scr(I); else return reg;
}
template <int I, typename T, typename IO> CLOUDSC_FN void park_st(IO &io, T &reg, T v) {
  if constexpr (io_scratch<IO>::value > I) io.scr(I) = v; else reg = v;
}

/* CLOUDSC_RELOAD (experiment, default off): with a staged IO policy the inputs of the level stay in shared memory until
 * level_end(), so the start-of-level values that are only needed again at the very end of the level (zqx0[0..3], zaorig)
 * can be rebuilt there from the staged rows (10 LDS + 5 FMA) instead of occupying 10 registers through the physics. */
#ifndef CLOUDSC_RELOAD
#define CLOUDSC_RELOAD 0
#endif

/* Optional level-range capability of an IO policy (`static constexpr bool kRanged`): the sweep of a column may be cut
 * into level ranges [k_begin, k_end) run at different times (the persistent kernel interleaves the bandwidth-bound upper
 * part of one column tile with the latency-bound cloudy part of another).  `Carry` -- and the parked slots -- is exactly
 * what has to travel across a cut: the policy provides  int k_begin(), k_end();  void load_carry(Carry<T>&),
 * save_carry(const Carry<T>&).  Every other policy sweeps [0, klev) and none of this is instantiated. */
template <typename IO, typename = void> struct io_ranged { static constexpr bool value = false; };
template <typename IO> struct io_ranged<IO, decltype((void)IO::kRanged)> { static constexpr bool value = IO::kRanged; };

/* State carried from level k-1 to level k (what "k-caching" keeps in registers) */
template <typename T>
struct Carry {
  T tp1_m1;      /* ztp1[jk-1]            cloudsc_c.c:1025,1868 */
  T a_m1;        /* za[jk-1] (clamped)    :1390,1597 */
  T pap_m1;      /* pap[jk-1]             :1026 */
  T paph_k;      /* paph[jk]              :803 */
  T mf_k;        /* pmfu[jk]+pmfd[jk]     :1005,1146 */
  T anewm1;      /* zanewm1               :2179 -> :1006,1027 */
  T qlnm1, qinm1;/* zqxnm1[l], zqxnm1[i]  :2419 -> :1013 */
  T pf_i, pf_r, pf_s; /* zpfplsx[i,r,s][jk] :2432 -> :1550 (liquid/vapour components are 0) */
  T covptot, covpmax; /* :1597-1608 */
  T cldtopdist;  /* :1390-1395 */
  T rainfrac;    /* prainfrac_toprfz :1875 -> :1881 */
  T fsqlf, fsqif, fcqlng, fcqnng, fsqltur, fsqitur; /* running flux sums :2548-2568 */
};

#ifndef CLOUDSC_ALFA_CLAMP
#define CLOUDSC_ALFA_CLAMP 1
#endif
/* FOEALFA / FOELDCPM (cloudsc_c.c:335-345, inlined by the reference at every use) */
template <typename T>
CLOUDSC_FN T foealfa(const CloudscConst<T> &c, T t) {
#if CLOUDSC_ALFA_CLAMP && !CLOUDSC_STRICT
  /* clamp the ramp instead of the temperature: both clamps on the integer pipe; differs from the reference form only for
     t >= rtwat, where (rtwat - rtice) * (1 / (rtwat - rtice)) may round to 1 - 1 ulp instead of 1 */
  const T x = CLOUDSC_MN1(CLOUDSC_MX0((t - c.rtice) * c.rtwat_rtice_r));
  return x * x;
#else
  const T x = (mx(c.rtice, mn(c.rtwat, t)) - c.rtice) * c.rtwat_rtice_r;
  return CLOUDSC_MN1R(x * x);
#endif
}
template <typename T>
CLOUDSC_FN T foeldcpm(const CloudscConst<T> &c, T al) { return al * c.ralvdcp + (T(1.0) - al) * c.ralsdcp; }

/* One Newton step of the saturation adjustment, cloudsc_c.c:1162-1168 (and :1169-1175). */
template <typename T>
CLOUDSC_FN void satadj_step(const CloudscConst<T> &c, T qp, T &tn, T &qsm) {
  const T one = T(1.0);
  const T al = foealfa(c, tn);
  const T dl = tn - c.r4les, di = tn - c.r4ies;
  const T rdl = CLOUDSC_RCP(dl), rdi = CLOUDSC_RCP(di);
  const T el = ex(CLOUDSC_DIVR(c.r3les * (tn - c.rtt), dl, rdl));
  const T ei = ex(CLOUDSC_DIVR(c.r3ies * (tn - c.rtt), di, rdi));
  T qsat = (c.r2es * (al * el + (one - al) * ei)) * qp;
  qsat = CLOUDSC_MNHR(qsat);
#if CLOUDSC_STRICT
  const T cor = one / (one - c.retv * qsat);
  qsat = qsat * cor;
  const T dem = (al * c.r5alvcp) * (one / (dl * dl)) + ((one - al) * c.r5alscp) * (one / (di * di)); /* FOEDEM */
  const T cond = (qsm - qsat) / (one + (qsat * cor) * dem);
#else
  /* (qsm - u/d) / (1 + (u/d)(1/d) dem) with d = 1 - retv u  ==  (qsm d - u) d / (d^2 + u dem):
     one reciprocal on the Newton step's critical path instead of two */
  const T dem = (al * c.r5alvcp) * (rdl * rdl) + ((one - al) * c.r5alscp) * (rdi * rdi);
  const T d = one - c.retv * qsat;
  const T cond = div_fast((qsm * d - qsat) * d, d * d + qsat * dem);
#endif
  tn = tn + foeldcpm(c, al) * cond;
  qsm = qsm - cond;
}

/*
 * The part of 3.7 (ice deposition, Rotstayn; cloudsc_c.c:1398-1431) that depends on the level's inputs only: the new ice
 * content zinew a pre-existing ice amount zice0 grows to, and zinfactor.  Everything downstream of it (zdepos and its
 * limits) depends on the first-guess liquid of this level.
 */
template <typename T>
CLOUDSC_FN void deposition_prep(const CloudscConst<T> &c, const T ptsphy, const T t, const T pap, const T rho, const T r_rho,
                                const T r_t, const T e_ice, const T fokoop, const T icecld, T &inew, T &ice0, T &infactor) {
  const T one = T(1.0);
  (void)r_rho; (void)r_t;
  const T vpice = CLOUDSC_DIVR((c.r2es * e_ice) * c.rv, c.rd, c.r_rd);
  const T vpliq = vpice * fokoop;
  const T icenuclei = T(1000.0) * CLOUDSC_EXP_COLD(CLOUDSC_DIV(c.p1296 * (vpliq - vpice), vpliq) - c.p0639);
#if CLOUDSC_STRICT
  const T add = (c.rlstt * (c.rlstt / (c.rv * t) - one)) / (c.p0024 * t);
  const T bdd = ((c.rv * t) * pap) / (c.p221 * vpice);
  const T cvds = ((c.p78 * pw(icenuclei / rho, c.p0666)) * (vpliq - vpice)) / ((c.p887 * (add + bdd)) * vpice);
  ice0 = mx(icecld, (icenuclei * c.riceinit) / rho);
#else
  const T add = (c.rlstt * ((c.rlstt * c.r_rv) * r_t - one)) * (c.r_p0024 * r_t);
  const T bdd = div_fast((c.rv * t) * pap, c.p221 * vpice);
  const T cvds = div_fast((c.p78 * CLOUDSC_POW(icenuclei * r_rho, c.p0666)) * (vpliq - vpice), (c.p887 * (add + bdd)) * vpice);
  ice0 = mx(icecld, (icenuclei * c.riceinit) * r_rho);
#endif
  inew = CLOUDSC_POW15((c.p0666 * cvds) * ptsphy + CLOUDSC_POW(ice0, c.p0666));
  infactor = CLOUDSC_MN1(CLOUDSC_DIVR(icenuclei, T(15000.0), c.r_15000));
}

/* CLOUDSC_DEPOS_EARLY (experiment): evaluate deposition_prep() next to the Newton steps of 3.4 -- both are long dependent
 * chains on the level's inputs only, in one basic block they overlap -- where supercooled liquid is already there at the
 * start of the level; the deposition block falls back to evaluating it itself where liquid only appears during the level. */
#ifndef CLOUDSC_DEPOS_EARLY
#define CLOUDSC_DEPOS_EARLY 0
#endif

/*
 * 5.2 sink limiter, cloudsc_c.c:2233-2286, on the antisymmetric representation of ZSOLQA (a_xy = zsolqa[x][y] for
 * x < y, diagonal d_m; species 0=l 1=i 2=r 3=s 4=v; the pair (i,s) and d_v are never written by any process).
 * (The first pass :2207-2227 is overwritten before use.)  zratio = zmm/max(zsinksum,zmm) is exactly 1 unless the sinks
 * of species m exceed its content, and scaling by 1 is the identity -- so the divide and the scaling only run in that
 * case.  Returns the column sums cs_m = sum_n zsolqa[n][m] of the limited matrix, i.e. the explicit terms of :2330.
 */
template <typename T>
CLOUDSC_FN void sink_limiter(const CloudscConst<T> &c, T &a01, T &a02, T &a03, T &a04, T &a12, T &a14, T &a23, T &a24, T &a34,
                             T &d0, T &d1, T &d2, T &d3, const T (&qx)[4], const T qv, T &cs0, T &cs1, T &cs2, T &cs3, T &cs4) {
  const T zero = T(0.0);
      /* column sums sum_n zsolqa[n][m], n = 0..4 in the reference's order (:2239, :2330) */
#if CLOUDSC_STRICT || defined(CLOUDSC_NOTRIM)
#define CLOUDSC_COLSUM0 ((((zero + d0) - a01) - a02) - a03) - a04
#define CLOUDSC_COLSUM1 (((zero + a01) + d1) - a12) - a14
#define CLOUDSC_COLSUM2 ((((zero + a02) + a12) + d2) - a23) - a24
#define CLOUDSC_COLSUM3 (((zero + a03) + a23) + d3) - a34
#define CLOUDSC_COLSUM4 (((zero + a04) + a14) + a24) + a34
#define CLOUDSC_NEG(x) (zero - (x))
#else
/* (0 + x and 0 - x as x and -x: the same values, signs of zeros aside, without the additions) */
#define CLOUDSC_COLSUM0 (((d0 - a01) - a02) - a03) - a04
#define CLOUDSC_COLSUM1 ((a01 + d1) - a12) - a14
#define CLOUDSC_COLSUM2 (((a02 + a12) + d2) - a23) - a24
#define CLOUDSC_COLSUM3 ((a03 + a23) + d3) - a34
#define CLOUDSC_COLSUM4 ((a04 + a14) + a24) + a34
#define CLOUDSC_NEG(x) (-(x))
#endif
      /* "if zsolqa[n][m] < 0 scale the pair": a_nm itself for n < m, a_mn > 0 for n > m; a negative
       * diagonal is scaled twice (n == m hits both statements of :2281-2282) */
#define CLOUDSC_SCALE_LO(a) if ((a) < zero) (a) = (a) * ratio
#define CLOUDSC_SCALE_HI(a) if ((a) > zero) (a) = (a) * ratio
#define CLOUDSC_SCALE_D(d) if ((d) < zero) (d) = ((d) * ratio) * ratio
      /* The five species are limited one after the other (a scaled pair feeds the later column sums),
       * but when no column exceeds its content -- the common case -- nothing is scaled at all and the
       * five unscaled sums are also the explicit sums of :2330.  So: all five sums and tests up front
       * (independent chains), and the sequential pass only if one of them fires; by induction over m
       * the pass cannot fire unless one of the up-front tests does. */
      cs0 = CLOUDSC_COLSUM0; cs1 = CLOUDSC_COLSUM1; cs2 = CLOUDSC_COLSUM2; cs3 = CLOUDSC_COLSUM3; cs4 = CLOUDSC_COLSUM4;
      const T mm0 = mx(qx[0], c.zepsec), mm1 = mx(qx[1], c.zepsec), mm2 = mx(qx[2], c.zepsec), mm3 = mx(qx[3], c.zepsec),
              mm4 = mx(qv, c.zepsec);
      if (CLOUDSC_RARE((CLOUDSC_NEG(cs0) > mm0) | (CLOUDSC_NEG(cs1) > mm1) | (CLOUDSC_NEG(cs2) > mm2) | (CLOUDSC_NEG(cs3) > mm3) | (CLOUDSC_NEG(cs4) > mm4))) {
        T sinksum = CLOUDSC_NEG(cs0);
        if (sinksum > mm0) {
          const T ratio = CLOUDSC_DIV(mm0, sinksum);
          CLOUDSC_SCALE_D(d0); CLOUDSC_SCALE_HI(a01); CLOUDSC_SCALE_HI(a02); CLOUDSC_SCALE_HI(a03); CLOUDSC_SCALE_HI(a04);
        }
        sinksum = CLOUDSC_NEG(CLOUDSC_COLSUM1);
        if (sinksum > mm1) {
          const T ratio = CLOUDSC_DIV(mm1, sinksum);
          CLOUDSC_SCALE_LO(a01); CLOUDSC_SCALE_D(d1); CLOUDSC_SCALE_HI(a12); CLOUDSC_SCALE_HI(a14);
        }
        sinksum = CLOUDSC_NEG(CLOUDSC_COLSUM2);
        if (sinksum > mm2) {
          const T ratio = CLOUDSC_DIV(mm2, sinksum);
          CLOUDSC_SCALE_LO(a02); CLOUDSC_SCALE_LO(a12); CLOUDSC_SCALE_D(d2); CLOUDSC_SCALE_HI(a23); CLOUDSC_SCALE_HI(a24);
        }
        sinksum = CLOUDSC_NEG(CLOUDSC_COLSUM3);
        if (sinksum > mm3) {
          const T ratio = CLOUDSC_DIV(mm3, sinksum);
          CLOUDSC_SCALE_LO(a03); CLOUDSC_SCALE_LO(a23); CLOUDSC_SCALE_D(d3); CLOUDSC_SCALE_HI(a34);
        }
        sinksum = CLOUDSC_NEG(CLOUDSC_COLSUM4);
        if (sinksum > mm4) {
          const T ratio = CLOUDSC_DIV(mm4, sinksum);
          CLOUDSC_SCALE_LO(a04); CLOUDSC_SCALE_LO(a14); CLOUDSC_SCALE_LO(a24); CLOUDSC_SCALE_LO(a34);
        }
        cs0 = CLOUDSC_COLSUM0; cs1 = CLOUDSC_COLSUM1; cs2 = CLOUDSC_COLSUM2; cs3 = CLOUDSC_COLSUM3; cs4 = CLOUDSC_COLSUM4;
      }
#undef CLOUDSC_COLSUM0
#undef CLOUDSC_COLSUM1
#undef CLOUDSC_COLSUM2
#undef CLOUDSC_COLSUM3
#undef CLOUDSC_COLSUM4
#undef CLOUDSC_SCALE_LO
#undef CLOUDSC_SCALE_HI
#undef CLOUDSC_SCALE_D
#undef CLOUDSC_NEG
}

/*
 * 5.2.2 implicit solve, cloudsc_c.c:2294-2412.  RHS :2326-2337 = qx + the limited explicit sums.  LHS: diag
 * den_m = 1 + fallsink_m + sum_n qb[m][n], off-diagonals -qb[m][n]; the only non-zero qb are (l,l) (i,i) (l,s) (i,s),
 * i.e. the transposed system is lower triangular with the single sub-diagonal row "snow <- liquid, ice".  The
 * reference's non-pivoting LU then produces exactly two multipliers, -qb03/den0 and -qb13/den1, leaves every diagonal
 * unchanged, and the back-substitution is a plain divide per species.  Then the clipping of :2402-2412.
 */
template <typename T>
CLOUDSC_FN void implicit_solve(const CloudscConst<T> &c, const T (&qx)[4], const T qv, const T cs0, const T cs1, const T cs2,
                               const T cs3, const T cs4, const T qb00, const T qb11, const T qb03, const T qb13,
                               const T fallsink1, const T fallsink2, const T fallsink3, T (&qxn)[5]) {
  const T zero = T(0.0), one = T(1.0);
      qxn[0] = qx[0] + cs0;
      qxn[1] = qx[1] + cs1;
      qxn[2] = qx[2] + cs2;
      qxn[3] = qx[3] + cs3;
      qxn[4] = qv + cs4;
      {
#if CLOUDSC_STRICT || defined(CLOUDSC_NOTRIM)
        const T den0 = ((one + zero) + qb00) + qb03;
#else
        const T den0 = (one + qb00) + qb03;
#endif
        const T den1 = ((one + fallsink1) + qb11) + qb13;
        const T den2 = one + fallsink2;
        const T den3 = one + fallsink3;
        const T r_den0 = CLOUDSC_RCP(den0), r_den1 = CLOUDSC_RCP(den1);
        (void)r_den0; (void)r_den1;
        const T l03 = CLOUDSC_DIVR(-qb03, den0, r_den0);
        const T l13 = CLOUDSC_DIVR(-qb13, den1, r_den1);
        qxn[3] = qxn[3] - l03 * qxn[0];
        qxn[3] = qxn[3] - l13 * qxn[1];
        qxn[3] = CLOUDSC_DIV(qxn[3], den3);
        qxn[2] = CLOUDSC_DIV(qxn[2], den2);
        qxn[1] = CLOUDSC_DIVR(qxn[1], den1, r_den1);
        qxn[0] = CLOUDSC_DIVR(qxn[0], den0, r_den0);
      }
      /* clip small/negative values into vapour, :2402-2412 -- as selects: the four tests are independent
       * and qxn[v] + 0 == qxn[v] where the reference adds nothing */
      {
        const bool c0 = qxn[0] < c.zepsec, c1 = qxn[1] < c.zepsec, c2 = qxn[2] < c.zepsec, c3 = qxn[3] < c.zepsec;
        qxn[4] = (((qxn[4] + (c0 ? qxn[0] : zero)) + (c1 ? qxn[1] : zero)) + (c2 ? qxn[2] : zero)) + (c3 ? qxn[3] : zero);
        qxn[0] = c0 ? zero : qxn[0]; qxn[1] = c1 ? zero : qxn[1]; qxn[2] = c2 ? zero : qxn[2]; qxn[3] = c3 ? zero : qxn[3];
      }
}

/*
 * The column sweep.  IO must provide
 *   template<int S> T ld(int k) const;        input slot S at level k (0-based), read while level k
 *                                             is being processed
 *   template<int S> T ld_next(int k) const;   row k+1 of slot S, read while level k is processed
 *                                             (IN_PAPH, IN_PMFU, IN_PMFD, IN_PLU only)
 *   T paph_row(int r) const;                  paph row r outside the level loop (r = 0, klev)
 *   template<int S> void st(int k, T v);      output slot S: row k (tendencies, pcovptot, plude) or
 *                                             row k+1 passed as k+1 (flux fields), once per level
 *   template<int S> void st_row0(T v);        row 0 of a flux field (before the level loop)
 *   T plsm() const; int ktype() const; void st_rainfrac(T);
 *   void level_begin(int k); void level_end(int k);   hooks for staged IO (may be empty)
 * and, only when instantiated with AER = true (yrecldp->laericesed / laericeauto, cloudsc_c.c:1561-1566,1625-1632):
 *   template<int A> T ld_aer(int k) const;    aerosol input A (AerSlot) at level k
 */
template <typename T, bool AER = false, typename IO>
CLOUDSC_FN void cloudsc_column(const CloudscConst<T> &c, const StepConst<T> &sc, IO &io, const int klev) {
  const T zero = T(0.0), one = T(1.0);
  const T ptsphy = sc.ptsphy;
  const T zqtmst = sc.zqtmst;
  constexpr bool kReload = (CLOUDSC_RELOAD != 0) && (io_scratch<IO>::value > 0);
#if !CLOUDSC_STRICT
  if (sizeof(T) == 8) exp2_tab_init();
#endif
  /* wet-bulb fit constants with exact binary64 high words, cloudsc_c.c:268,270 */
  const T ztw3 = T(0.85e5), ztw5 = T(275.0);

  const int ktop = c.ncldtop - 1; /* first level (0-based) of the physics sweep, :732 */

  /* per-column constants */
  T paph_surf_r = zero, r_paph_surf_r = zero;   /* registers, unless parked */
  {
    const T ps = io.paph_row(klev); /* paph[klev+1] :1260,2064 */
    park_st<PARK_PAPH_SURF>(io, paph_surf_r, ps);
    park_st<PARK_R_PAPH_SURF>(io, r_paph_surf_r, CLOUDSC_RCP(ps));
  }
  const T plsm = io.plsm();
  const int ktype = io.ktype();
  /* cloudsc_c.c:1710-1717 */
  const bool land = plsm > T(0.5);
#define paph_surf park_ld<PARK_PAPH_SURF>(io, paph_surf_r)
#define r_paph_surf park_ld<PARK_R_PAPH_SURF>(io, r_paph_surf_r)
#define kk_num_pow (land ? c.kk_num_pow_land : c.kk_num_pow_sea)
#define lcrit_kk (land ? c.rclcrit_land : c.rclcrit_sea)
  (void)r_paph_surf_r;

  int k_begin = 0, k_end = klev;
  if constexpr (io_ranged<IO>::value) { k_begin = io.k_begin(); k_end = io.k_end(); }
  Carry<T> s;
  if (k_begin == 0) {
  s.tp1_m1 = zero; s.a_m1 = zero; s.pap_m1 = zero;
  s.paph_k = io.paph_row(0);
  s.mf_k = zero;
  s.anewm1 = zero; s.qlnm1 = zero; s.qinm1 = zero;           /* :718 */
  s.pf_i = zero; s.pf_r = zero; s.pf_s = zero;               /* :493 */
  s.covptot = zero; s.covpmax = zero; s.cldtopdist = zero;   /* :720-723 */
  s.rainfrac = zero;                                         /* :512 */
  /* (parked by IO policies with scratch: loaded where the level first needs them, stored after their last update) */
  park_st<PARK_PF_I>(io, s.pf_i, zero); park_st<PARK_PF_R>(io, s.pf_r, zero); park_st<PARK_PF_S>(io, s.pf_s, zero);
  park_st<PARK_COVPTOT>(io, s.covptot, zero); park_st<PARK_COVPMAX>(io, s.covpmax, zero);
  park_st<PARK_CLDTOPDIST>(io, s.cldtopdist, zero); park_st<PARK_RAINFRAC>(io, s.rainfrac, zero);
  park_st<PARK_FSQLF>(io, s.fsqlf, zero); park_st<PARK_FSQIF>(io, s.fsqif, zero); park_st<PARK_FCQLNG>(io, s.fcqlng, zero);
  park_st<PARK_FCQNNG>(io, s.fcqnng, zero); park_st<PARK_FSQLTUR>(io, s.fsqltur, zero); park_st<PARK_FSQITUR>(io, s.fsqitur, zero);

  /* row 0 of every flux output, :2533-2542 and zpfplsx[.][0]=0 -> :2523-2524,2578-2579 */
  io.template st_row0<OUT_FSQLF>(zero);  io.template st_row0<OUT_FSQIF>(zero);
  io.template st_row0<OUT_FSQRF>(zero);  io.template st_row0<OUT_FSQSF>(zero);
  io.template st_row0<OUT_FCQLNG>(zero); io.template st_row0<OUT_FCQNNG>(zero);
  io.template st_row0<OUT_FCQRNG>(zero); io.template st_row0<OUT_FCQSNG>(zero);
  io.template st_row0<OUT_FSQLTUR>(zero); io.template st_row0<OUT_FSQITUR>(zero);
  io.template st_row0<OUT_FPLSL>(zero);  io.template st_row0<OUT_FPLSN>(zero);
  io.template st_row0<OUT_FHPSL>(-c.rlvtt * zero); io.template st_row0<OUT_FHPSN>(-c.rlstt * zero);
  } else {
    if constexpr (io_ranged<IO>::value) io.load_carry(s);   /* the sweep resumes below a cut */
  }

  for (int k = k_begin; k < k_end; ++k) {
    io.level_begin(k);
    /* ================= 1. initial values of this level, :462-485 ========================== */
    const T pap = io.template ld<IN_PAP>(k);
    const T paph_n = io.template ld_next<IN_PAPH>(k);
    const T t = io.template ld<IN_PT>(k) + ptsphy * io.template ld<IN_TT>(k);
    T qv = io.template ld<IN_PQ>(k) + ptsphy * io.template ld<IN_TQ>(k);
    T a = io.template ld<IN_PA>(k) + ptsphy * io.template ld<IN_TA>(k);
    const T aorig = a;
    T qx[4], qx0[4];
    qx[0] = io.template ld<IN_PCLV0>(k) + ptsphy * io.template ld<IN_TCLD0>(k);
    qx[1] = io.template ld<IN_PCLV1>(k) + ptsphy * io.template ld<IN_TCLD1>(k);
    qx[2] = io.template ld<IN_PCLV2>(k) + ptsphy * io.template ld<IN_TCLD2>(k);
    qx[3] = io.template ld<IN_PCLV3>(k) + ptsphy * io.template ld<IN_TCLD3>(k);
#pragma unroll
    for (int m = 0; m < 4; ++m) qx0[m] = qx[m];
    T tend_t = zero, tend_q = zero;

    /* ---- tidy up very small cloud cover / condensate, :519-541, and small CLV variables, :547-575
     * (iphase = l:1 i:2 r:1 s:2).  Written as selects: where the reference does nothing, a zero is
     * added (x + 0 == x), so the values are the reference's and the block is branch-free -- it shares
     * one basic block with the saturation chains below, whose latencies it helps to fill. */
    {
      const bool clr = (qx[0] + qx[1] < c.rlmin) | (a < c.ramin);
      T v = clr ? qx[0] : zero;
      T qadj = v * zqtmst;
      tend_q = tend_q + qadj;
      tend_t = tend_t - c.ralvdcp * qadj;
      qv = qv + v;
      qx[0] = clr ? zero : qx[0];
      v = clr ? qx[1] : zero;
      qadj = v * zqtmst;
      tend_q = tend_q + qadj;
      tend_t = tend_t - c.ralsdcp * qadj;
      qv = qv + v;
      qx[1] = clr ? zero : qx[1];
      a = clr ? zero : a;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const bool small = qx[m] < c.rlmin;
        v = small ? qx[m] : zero;
        qadj = v * zqtmst;
        tend_q = tend_q + qadj;
        tend_t = tend_t - ((m & 1) ? c.ralsdcp : c.ralvdcp) * qadj;
        qv = qv + v;
        qx[m] = small ? zero : qx[m];
      }
    }

    /* ---- FOEALFA of this level (:588; also needed by the flux section at every level) */
    const T alfa_w = foealfa(c, t);
    /* ---- cloud fraction in [0,1], :625 */
    a = CLOUDSC_MX0R(CLOUDSC_MN1R(a));

    /* values of this level the flux section needs (zqxn2d stays 0 above ncldtop, :503) */
    T qxn2d[4] = {zero, zero, zero, zero};
    T plude_out = io.template ld<IN_PLUDE>(k); /* plude[jk] as the flux section sees it, :2559 */
    T tend_a = zero, tcld[4] = {zero, zero, zero, zero}, covptot_out = zero;
    T pf_i_n = zero, pf_r_n = zero, pf_s_n = zero;  /* zpfplsx[.][jk+1] */
    /* pmfu+pmfd of the level below (:1067,1143); becomes mf_k of the next iteration (needs ncldtop >= 2) */
    const T mf_n = (k + 1 < klev && k + 1 >= ktop) ? io.template ld_next<IN_PMFU>(k) + io.template ld_next<IN_PMFD>(k) : zero;

    if (k >= ktop) {
      /* ================= saturation values of this level, :583-640 ======================== */
      /* (the reference computes them for all levels; above ncldtop only zfoealfa is ever read) */
      const T dl = t - c.r4les, di = t - c.r4ies;
      const T r_dl = CLOUDSC_RCP(dl), r_di = CLOUDSC_RCP(di);
      const T r_pap = CLOUDSC_RCP(pap);
      const T e_liq = ex(CLOUDSC_DIVR(c.r3les * (t - c.rtt), dl, r_dl));
      const T e_ice = ex(CLOUDSC_DIVR(c.r3ies * (t - c.rtt), di, r_di));
      const T foeewmt = CLOUDSC_MNH(CLOUDSC_DIVR(c.r2es * (alfa_w * e_liq + (one - alfa_w) * e_ice), pap, r_pap));
      const T cor_m = CLOUDSC_RCP(one - c.retv * foeewmt);
      const T qsmix = CLOUDSC_DIVR(foeewmt, one - c.retv * foeewmt, cor_m);
      const T zalfa = CLOUDSC_MX0R(cs(one, t - c.rtt));
      /* (the reference applies min(0.5, .) twice here; the second one is the identity) */
      const T foeew = CLOUDSC_MNH(CLOUDSC_DIVR(zalfa * (c.r2es * e_liq) + (one - zalfa) * (c.r2es * e_ice), pap, r_pap));
      const T cor_i = CLOUDSC_RCP(one - c.retv * foeew);
      const T qsice = CLOUDSC_DIVR(foeew, one - c.retv * foeew, cor_i);
      /* zqsliq (:607-609) is only read by the rain-evaporation block and is evaluated there */
      /* liquid/ice fractions, :628-636 */
      const T li = qx[0] + qx[1];
      const bool has_li = li > c.rlmin;
      const T liqfrac_u = CLOUDSC_DIV(qx[0], li);      /* discarded (possibly inf/nan) where li <= rlmin */
      const T liqfrac = has_li ? liqfrac_u : zero;
      const T icefrac = has_li ? one - liqfrac_u : zero;

      /* ================= 3.0 initialise, :739-859 ========================================= */
      T qxfg[4] = {qx[0], qx[1], qx[2], qx[3]};
      /* ZSOLQA (:742-750).  Every process of the scheme enters it as an antisymmetric pair
       * zsolqa[x][y] += v; zsolqa[y][x] -= v (a transfer between two species) or as a diagonal source,
       * and the sink limiter scales both members of a pair by the same factor, so zsolqa[y][x] ==
       * -zsolqa[x][y] holds bit for bit (IEEE addition and multiplication are sign-symmetric).  Only
       * the upper triangle a_xy = zsolqa[x][y] (x < y; 0=l 1=i 2=r 3=s 4=v; the pair (i,s) is never
       * written) and the diagonal d_m are kept: 13 instead of 25 values in registers. */
      T a01 = zero, a02 = zero, a03 = zero, a04 = zero, a12 = zero, a14 = zero, a23 = zero, a24 = zero, a34 = zero;
      T d0 = zero, d1 = zero, d2 = zero, d3 = zero;
      T qb03 = zero, qb13 = zero, qb00 = zero, qb11 = zero;
      T solab = zero, solac = zero;
      T convsrce0 = zero, convsrce1 = zero, psupsrce0 = zero, psupsrce1 = zero;
      T convsink = zero; /* zconvsink[l] == zconvsink[i] */

      const T dp = paph_n - s.paph_k;
      const T r_dp = CLOUDSC_RCP(dp);
      const T gdp = CLOUDSC_DIVR(c.rg, dp, r_dp);
#if CLOUDSC_STRICT
      const T rho = pap / (c.rd * t);
      const T r_rho = one / rho;
      const T r_t = one / t;   /* (unused in strict mode: deposition_prep() divides) */
#else
      const T r_t = rcp_fast(t);
      const T rho = (pap * c.r_rd) * r_t;
      const T r_rho = (c.rd * t) * r_pap;
#endif
      (void)r_rho;
      const T dtgdp = ptsphy * gdp;
      const T rdtgdp = dp * sc.r_ptsphy_rg;
      /* dqs/dT correction factors, :818-834 (zcorqsliq/zevaplimliq/zevaplimice are never read) */
#if CLOUDSC_STRICT
      const T facw = c.r5les / (dl * dl);
      const T faci = c.r5ies / (di * di);
#else
      const T facw = c.r5les * (r_dl * r_dl);
      const T faci = c.r5ies * (r_di * r_di);
#endif
      const T dqsicedt = (faci * cor_i) * qsice;
      const T corqsice = one + c.ralsdcp * dqsicedt;
      const T r_corqsice = CLOUDSC_RCP(corqsice);
      (void)r_corqsice;
      const T facm = alfa_w * facw + (one - alfa_w) * faci;
      const T dqsmixdt = (facm * cor_m) * qsmix;
      const T corqsmix = one + foeldcpm(c, alfa_w) * dqsmixdt;
      const T evaplimmix = CLOUDSC_MX0(CLOUDSC_DIV(qsmix - qv, corqsmix));
      /* in-cloud condensate, :837-840 */
      const T tmpa = CLOUDSC_RCP(mx(a, c.zepsec));
      T liqcld = qx[0] * tmpa;
      T icecld = qx[1] * tmpa;
      const T licld = liqcld + icecld;
      /* 1/max(eps, 1-a): :1278,1284,1321,1940,1990,2051,2053 */
      const T one_m_a = one - a;
      const T r_1ma = CLOUDSC_RCP(mx(c.zepsec, one_m_a));
      (void)r_1ma;

      const T fokoop = mn(c.rkoop1 - c.rkoop2 * t, CLOUDSC_DIV(c.r2es * e_liq, c.r2es * e_ice));
      /* ---- does anything read zdqs at this point?  3.4a / 3.4b(1) act on existing cloud only (a > 0; with
       * a == 0 they add zeros), 3.4b(2) only inside the humidity window below, which does not depend on
       * zdqs.  Where neither holds -- every clear, dry point: the whole stratosphere -- the two Newton
       * steps (4 exp, 6 reciprocals) are not evaluated at all.  Exact: zdqs is dead there. */
      T rhc = c.ramid;
      {
        const T sigk = CLOUDSC_DIVR(pap, paph_surf, r_paph_surf);
        if (sigk > c.p08) { const T x = CLOUDSC_DIVR(sigk - c.p08, c.p02, c.r_p02); rhc = c.ramid + sc.one_m_ramid * (x * x); }
      }
      T qe_nc;                                         /* zqe of :1275-1302 */
      if (c.nssopt == 0 || c.nssopt == 1) {
        qe_nc = CLOUDSC_DIVR(qv - a * qsice, mx(c.zepsec, one_m_a), r_1ma);
        qe_nc = CLOUDSC_MX0R(qe_nc);
      } else if (c.nssopt == 2) {
        qe_nc = qv;
      } else {
        qe_nc = qv + li;
      }
      const T fac_nc = (c.nssopt == 0 || t >= c.rtt) ? one : fokoop;
      const bool new_cloud_window = a < sc.one_m_zepsec && qe_nc >= qsice * fac_nc * rhc && qe_nc < qsice * fac_nc;
      const bool need_dqs = (a > zero) | new_cloud_window;
      /* ================= 3.4 condensation/evaporation due to dqsat/dT, :1137-1182 ========= */
      /* (evaluated here, ahead of 3.1-3.4: it is a pure function of this level's inputs, and next to the
       * saturation block its two Newton steps overlap with that block's exp/reciprocal chains) */
      T dqs = zero;
      T dep_inew = zero, dep_ice0 = zero, dep_infactor = zero;
      bool dep_ready = false;
      (void)dep_ready;
      if (need_dqs) {
        const T dtdp = CLOUDSC_DIVR(sc.zrdcp * t, pap, r_pap);
        const T dpmxdt = dp * zqtmst;
        const T mfdn = (k + 1 < klev) ? mf_n : zero;
        T wtot = io.template ld<IN_PVERVEL>(k) + sc.half_rg * (s.mf_k + mfdn);
        wtot = mn(dpmxdt, mx(-dpmxdt, wtot));
        const T zzzdt = io.template ld<IN_PHRSW>(k) + io.template ld<IN_PHRLW>(k);
        const T dtdiab = mn(dpmxdt * dtdp, mx(-dpmxdt * dtdp, zzzdt)) * ptsphy; /* + ralfdcp*zldefr, zldefr == 0 */
        const T dtforc = (dtdp * wtot) * ptsphy + dtdiab;
        T tn = t + dtforc;
        tn = CLOUDSC_MX160(tn);
        T qsm = qsmix;
#if CLOUDSC_DEPOS_EARLY && !CLOUDSC_STRICT
#if CLOUDSC_DEPOS_EARLY == 1
        if (qx[0] > c.rlmin && t < c.rtt)
#endif
        { /* (== 2: for every lane that is here, without a branch -- the values are only read where 3.7 applies) */
          deposition_prep<T>(c, ptsphy, t, pap, rho, r_rho, r_t, e_ice, fokoop, icecld, dep_inew, dep_ice0, dep_infactor);
          dep_ready = true;
        }
#endif
        satadj_step(c, r_pap, tn, qsm);
        satadj_step(c, r_pap, tn, qsm);
        dqs = qsm - qsmix;
      }
      /* evaporate very small amounts of liquid and ice, :847-857 */
      a04 = qx[0] < c.rlmin ? qx[0] : zero;
      a14 = qx[1] < c.rlmin ? qx[1] : zero;

      /* ================= 3.1 ice supersaturation adjustment, :879-952 ===================== */
      {
        T fac, facI;
        if (c.nssopt == 0 || t >= c.rtt) { fac = one; facI = one; }
        else { fac = a + fokoop * one_m_a; facI = sc.faci_koop; }
        T supsat;
        if (a > sc.one_m_ramin) {
          supsat = CLOUDSC_MX0(CLOUDSC_DIVR(qv - fac * qsice, corqsice, r_corqsice));
        } else {
          const T qp1env = CLOUDSC_DIV(qv - a * qsice, mx(one_m_a, c.zepsilon));
          supsat = CLOUDSC_MX0(CLOUDSC_DIVR(one_m_a * (qp1env - fac * qsice), corqsice, r_corqsice));
        }
        if (CLOUDSC_RARE(supsat > c.zepsec)) {
          if (t > c.rthomo) { a04 = a04 - supsat; qxfg[0] = qxfg[0] + supsat; }
          else              { a14 = a14 - supsat; qxfg[1] = qxfg[1] + supsat; }
          solac = one_m_a * facI;
        }
        const T psupsat = io.template ld<IN_PSUPSAT>(k);
        if (CLOUDSC_RARE(psupsat > c.zepsec)) {
          if (t > c.rthomo) { d0 = d0 + psupsat; psupsrce0 = psupsat; qxfg[0] = qxfg[0] + psupsat; }
          else              { d1 = d1 + psupsat; psupsrce1 = psupsat; qxfg[1] = qxfg[1] + psupsat; }
          solac = one_m_a * facI;
        }
      }

      /* ================= 3.2 detrainment from convection, :967-987 ======================== */
      if (k + 1 < klev) {
        const T plu_n = io.template ld_next<IN_PLU>(k);
        plude_out = plude_out * dtgdp;
        if (CLOUDSC_RARE(plu_n > c.zepsec && plude_out > c.rlmin)) {
          solac = solac + CLOUDSC_DIV(plude_out, plu_n);
          convsrce0 = alfa_w * plude_out;
          convsrce1 = (one - alfa_w) * plude_out;
          d0 = d0 + convsrce0;
          d1 = d1 + convsrce1;
        } else {
          plude_out = zero;
        }
        d3 = d3 + io.template ld<IN_PSNDE>(k) * dtgdp;
      }

      /* ================= 3.3 subsidence compensating convective updraughts, :1002-1075 ==== */
      if (k > ktop) {
        const T mf = CLOUDSC_MX0R(s.mf_k * dtgdp);
        T acust = mf * s.anewm1;
        const T lcust0 = mf * s.qlnm1, lcust1 = mf * s.qinm1;
        convsrce0 = convsrce0 + lcust0;
        convsrce1 = convsrce1 + lcust1;
        const T dtdp = CLOUDSC_DIV(sc.zrdcp_half * (s.tp1_m1 + t), s.paph_k);
        const T dtforc = dtdp * (pap - s.pap_m1);
        const T dqs_s = (s.anewm1 * dtforc) * dqsmixdt;
        T lfinalsum = zero;
        {
          T lfinal = CLOUDSC_MX0R(lcust0 - dqs_s);
          T evap = mn(lcust0 - lfinal, evaplimmix);
          lfinal = lcust0 - evap;
          lfinalsum = lfinalsum + lfinal;
          d0 = d0 + lcust0; a04 = a04 + evap;
          lfinal = CLOUDSC_MX0R(lcust1 - dqs_s);
          evap = mn(lcust1 - lfinal, evaplimmix);
          lfinal = lcust1 - evap;
          lfinalsum = lfinalsum + lfinal;
          d1 = d1 + lcust1; a14 = a14 + evap;
        }
        if (lfinalsum < c.zepsec) acust = zero;
        solac = solac + acust;
      }
      if (k + 1 < klev) { /* subsidence sink, :1064-1075 */
        const T mfdn = CLOUDSC_MX0R(mf_n * dtgdp);
        solab = solab + mfdn;
        qb00 = qb00 + mfdn;
        qb11 = qb11 + mfdn;
        convsink = mfdn;
      }

      /* ================= 3.4 erosion of clouds by turbulent mixing, :1087-1118 ============ */
      {
        const T ldifdt = (ktype > 0 && plude_out > c.zepsec) ? sc.ldifdt_conv : sc.ldifdt;
        if (li > c.zepsec) {
          const T e = ldifdt * CLOUDSC_MX0(qsmix - qv);
          T leros = a * e;
          leros = mn(leros, evaplimmix);
          leros = mn(leros, li);
          const T aeros = CLOUDSC_DIV(leros, licld);
          solac = solac - aeros;
          a04 = a04 + liqfrac * leros;
          a14 = a14 + icefrac * leros;
        }
      }

      /* 3.4a evaporation of clouds, :1191-1205 */
      if (dqs > zero) {
        T levap = a * mn(dqs, licld);
        levap = mn(levap, evaplimmix);
        levap = mn(levap, CLOUDSC_MX0(qsmix - qv));
        a04 = a04 + liqfrac * levap;
        a14 = a14 + icefrac * levap;
      }
      /* 3.4b (1) increase of cloud water in existing clouds, :1214-1248 */
      if (dqs <= -c.rlmin && a > c.zepsec) {
        T lcond1 = CLOUDSC_MX0(-dqs);
        T cdmax;
        if (a > c.p099) {
          const T corm = CLOUDSC_RCP(one - c.retv * qsmix);
#if CLOUDSC_STRICT
          const T dem = (alfa_w * c.r5alvcp) * (one / (dl * dl)) + ((one - alfa_w) * c.r5alscp) * (one / (di * di));
#else
          const T dem = (alfa_w * c.r5alvcp) * (r_dl * r_dl) + ((one - alfa_w) * c.r5alscp) * (r_di * r_di);
#endif
          cdmax = CLOUDSC_DIV(qv - qsmix, one + (corm * qsmix) * dem);
        } else {
          cdmax = CLOUDSC_DIV(qv - a * qsmix, a);
        }
        lcond1 = CLOUDSC_MX0(mn(lcond1, cdmax));
        lcond1 = a * lcond1;
        if (lcond1 < c.rlmin) lcond1 = zero;
        if (t > c.rthomo) { a04 = a04 - lcond1; qxfg[0] = qxfg[0] + lcond1; }
        else              { a14 = a14 - lcond1; qxfg[1] = qxfg[1] + lcond1; }
      }
      /* 3.4b (2) generation of new clouds, :1254-1361 */
      if (dqs <= -c.rlmin && new_cloud_window) {
        const T qe = qe_nc, fac = fac_nc;
        {
          T acond = CLOUDSC_DIV(-(one_m_a * fac) * dqs, mx(T(2.0) * (fac * qsice - qe), c.zepsec));
          acond = mn(acond, one_m_a);
          T lcond2 = (-(fac * dqs) * T(0.5)) * acond;
          const T zdl = CLOUDSC_DIVR(T(2.0) * (fac * qsice - qe), mx(c.zepsec, one_m_a), r_1ma);
          if (dqs * fac < -zdl) {
            const T lcondlim = ((a - one) * fac) * dqs - fac * qsice + qv;
            lcond2 = mn(lcond2, lcondlim);
          }
          lcond2 = CLOUDSC_MX0(lcond2);
          if (one_m_a < c.zepsec || lcond2 < c.rlmin) { lcond2 = zero; acond = zero; }
          if (lcond2 == zero) acond = zero;
          solac = solac + acond;
          if (t > c.rthomo) { a04 = a04 - lcond2; qxfg[0] = qxfg[0] + lcond2; }
          else              { a14 = a14 - lcond2; qxfg[1] = qxfg[1] + lcond2; }
        }
      }

      /* ================= 3.7 growth of ice by vapour deposition (Rotstayn), :1384-1447 ==== */
      s.cldtopdist = park_ld<PARK_CLDTOPDIST>(io, s.cldtopdist);
      if (a >= c.rcldtopcf && s.a_m1 < c.rcldtopcf) s.cldtopdist = zero;
#if CLOUDSC_STRICT
      else s.cldtopdist = s.cldtopdist + dp / (rho * c.rg);
#else
      else s.cldtopdist = s.cldtopdist + (dp * r_rho) * sc.zrg_r;
#endif
      park_st<PARK_CLDTOPDIST>(io, s.cldtopdist, s.cldtopdist);
      if (qxfg[0] > c.rlmin && t < c.rtt) {
        if (!dep_ready) deposition_prep<T>(c, ptsphy, t, pap, rho, r_rho, r_t, e_ice, fokoop, icecld, dep_inew, dep_ice0, dep_infactor);
        const T inew = dep_inew, ice0 = dep_ice0, infactor = dep_infactor;
        T depos = CLOUDSC_MX0(a * (inew - ice0));
        depos = mn(depos, qxfg[0]);
        depos = depos * CLOUDSC_MN1(infactor + (one - infactor) * (c.rdepliqrefrate + CLOUDSC_DIVR(s.cldtopdist, c.rdepliqrefdepth, c.r_rdepliqrefdepth)));
        a01 = a01 + depos;
        qxfg[1] = qxfg[1] + depos; qxfg[0] = qxfg[0] - depos;
      }

      /* ================= 4. precipitation processes ====================================== */
      /* revised in-cloud condensate :1529-1532 (same 1/max(a,eps) as :837) */
      liqcld = qxfg[0] * tmpa;
      icecld = qxfg[1] * tmpa;

      /* 4.2 sedimentation, :1541-1576 (species i, r, s; llfall = r,s plus ice) */
      T qpretot = zero;
      T fallsrce1 = zero, fallsrce2 = zero, fallsrce3 = zero;
      if (k > ktop) {
        fallsrce1 = park_ld<PARK_PF_I>(io, s.pf_i) * dtgdp; d1 = d1 + fallsrce1; qxfg[1] = qxfg[1] + fallsrce1; qpretot = qpretot + qxfg[1];
        fallsrce2 = park_ld<PARK_PF_R>(io, s.pf_r) * dtgdp; d2 = d2 + fallsrce2; qxfg[2] = qxfg[2] + fallsrce2; qpretot = qpretot + qxfg[2];
        fallsrce3 = park_ld<PARK_PF_S>(io, s.pf_s) * dtgdp; d3 = d3 + fallsrce3; qxfg[3] = qxfg[3] + fallsrce3; qpretot = qpretot + qxfg[3];
      }
      /* aerosol override of the ice fall speed, :1561-1566: zvqx[ice] = 0.002*pow(pre_ice, 1.0) (pow(x, 1) == x).  The
       * reference overwrites the shared zvqx[1] inside its jl loop, but reads it back for the same jl on the next line
       * and nowhere else during the sweep (:447 runs before it), so nothing leaks between columns. */
      T vice = c.rvice;
      if constexpr (AER) { if (c.laericesed) vice = c.p0002 * io.template ld_aer<AER_RE_ICE>(k); }
      const T fallsink1 = dtgdp * (vice * rho);
      const T fallsink2 = dtgdp * (c.rvrain * rho);
      const T fallsink3 = dtgdp * (c.rvsnow * rho);

      /* precipitation cover, max-random overlap, :1595-1609 */
      T covpclr, raincld, snowcld;
      s.covptot = park_ld<PARK_COVPTOT>(io, s.covptot);
      s.covpmax = park_ld<PARK_COVPMAX>(io, s.covpmax);
      if (qpretot > c.zepsec) {
        s.covptot = one - CLOUDSC_DIV((one - s.covptot) * (one - mx(a, s.a_m1)), one - mn(s.a_m1, c.one_m_1em6));
        s.covptot = mx(s.covptot, c.rcovpmin);
        covpclr = CLOUDSC_MX0R(s.covptot - a);
        const T r_cov = CLOUDSC_RCP(s.covptot);
        (void)r_cov;
        raincld = CLOUDSC_DIVR(qxfg[2], s.covptot, r_cov);
        snowcld = CLOUDSC_DIVR(qxfg[3], s.covptot, r_cov);
        s.covpmax = mx(s.covptot, s.covpmax);
      } else {
        raincld = zero; snowcld = zero; s.covptot = zero; covpclr = zero; s.covpmax = zero;
      }

      /* 4.3a autoconversion to snow, :1617-1637 */
      if (t <= c.rtt) {
        if (icecld > c.zepsec) {
          T zco = sc.snowlin * CLOUDSC_EXP_COLD(c.rsnowlin2 * (t - c.rtt));
          T x = CLOUDSC_DIVR(icecld, c.rlcritsnow, c.r_rlcritsnow);
          if constexpr (AER) {
            if (c.laericeauto) {   /* :1625-1632: critical ice content and rate scaled by the ice-nuclei concentration */
              zco = zco * CLOUDSC_POW(CLOUDSC_DIV(c.rnice, io.template ld_aer<AER_NICE>(k)), c.p0333);
              x = CLOUDSC_DIV(icecld, io.template ld_aer<AER_ICRIT>(k));
            }
          }
          const T snowaut = zco * (one - CLOUDSC_EXP_LO_COLD(-(x * x)));
          qb13 = qb13 + snowaut;
        }
      }
      /* 4.3b warm-rain autoconversion + accretion, Khairoutdinov-Kogan, :1644-1759 */
      if (liqcld > c.zepsec) {
        T rainaut, rainacc;
        if (CLOUDSC_RARE(liqcld > lcrit_kk)) {
          rainaut = ((((T(1.5) * a) * ptsphy) * c.rcl_kkaau) * CLOUDSC_POW(liqcld, c.rcl_kkbauq)) * kk_num_pow;
          rainaut = mn(rainaut, qxfg[0]);
          if (rainaut < c.zepsec) rainaut = zero;
          rainacc = (((T(2.0) * a) * ptsphy) * c.rcl_kkaac) * CLOUDSC_POW(liqcld * raincld, c.rcl_kkbac);
          rainacc = mn(rainacc, qxfg[0]);
          if (rainacc < c.zepsec) rainacc = zero;
        } else {
          rainaut = zero; rainacc = zero;
        }
        if (t <= c.rtt) {
          a03 = a03 + rainaut; a03 = a03 + rainacc;
        } else {
          a02 = a02 + rainaut; a02 = a02 + rainacc;
        }
      }
      /* riming of snow by cloud water, :1771-1785 */
      if (t <= c.rtt && liqcld > c.zepsec) {
        if (CLOUDSC_RARE(s.covptot > c.p001 && snowcld > c.zepsec)) {
          const T fallcorr = CLOUDSC_POW(CLOUDSC_DIVR(c.rdensref, rho, r_rho), c.p04);
          T snowrime = ((((c.p03 * s.covptot) * ptsphy) * c.rcl_const7s) * fallcorr) *
                       CLOUDSC_POW((rho * snowcld) * c.rcl_const1s, c.rcl_const8s);
          snowrime = CLOUDSC_MN1(snowrime);
          qb03 = qb03 + snowrime;
        }
      }

      /* 4.4a melting of snow and ice, :1817-1859 (imelt: ice->rain, snow->rain) */
      {
        const T icetot = qxfg[1] + qxfg[3];
        T meltmax = zero;
        if (CLOUDSC_RARE(icetot > c.zepsec && t > c.rtt)) {
          const T subsat = CLOUDSC_MX0(qsice - qv);
          const T tdmtw0 = t - c.rtt - subsat * (c.ztw1 + c.ztw2 * (pap - ztw3) - c.ztw4 * (t - ztw5));
          const T cons1 = ab(CLOUDSC_DIVR(ptsphy * (one + T(0.5) * tdmtw0), c.rtaumel, c.r_rtaumel));
          meltmax = CLOUDSC_MX0((tdmtw0 * cons1) * sc.zrldcp);
        }
        if (CLOUDSC_RARE(icetot > c.zepsec && meltmax > c.zepsec)) {
          const T r_icetot = CLOUDSC_RCP(icetot);
          (void)r_icetot;
          T alfa = CLOUDSC_DIVR(qxfg[1], icetot, r_icetot);
          T melt = mn(qxfg[1], alfa * meltmax);
          qxfg[1] = qxfg[1] - melt; qxfg[2] = qxfg[2] + melt;
          a12 = a12 + melt;
          alfa = CLOUDSC_DIVR(qxfg[3], icetot, r_icetot);
          melt = mn(qxfg[3], alfa * meltmax);
          qxfg[3] = qxfg[3] - melt; qxfg[2] = qxfg[2] + melt;
          a23 = a23 - melt;
        }
      }
      /* 4.4b freezing of rain, :1866-1906 */
      if (CLOUDSC_RARE(qx[2] > c.zepsec)) {
        if (t <= c.rtt && s.tp1_m1 > c.rtt) {
          const T qpre = mx(qx[3] + qx[2], c.zepsec);
          park_st<PARK_RAINFRAC>(io, s.rainfrac, CLOUDSC_DIV(qx[2], qpre));
        }
        if (t < c.rtt) {
          T frzmax;
          if (park_ld<PARK_RAINFRAC>(io, s.rainfrac) > c.p08) {
            const T lambda = CLOUDSC_POW(CLOUDSC_DIV(c.rcl_fac1, rho * qx[2]), c.rcl_fac2);
            const T temp = c.rcl_fzrab * (t - c.rtt);
            const T frz = ((ptsphy * CLOUDSC_DIVR(c.rcl_const5r, rho, r_rho)) * (CLOUDSC_EXP_COLD(temp) - one)) * CLOUDSC_POW(lambda, c.rcl_const6r);
            frzmax = CLOUDSC_MX0(frz);
          } else {
            const T cons1 = ab(CLOUDSC_DIVR(ptsphy * (one + T(0.5) * (c.rtt - t)), c.rtaumel, c.r_rtaumel));
            frzmax = CLOUDSC_MX0(((c.rtt - t) * cons1) * sc.zrldcp);
          }
          if (frzmax > c.zepsec) {
            const T frz = mn(qx[2], frzmax);
            a23 = a23 + frz;
          }
        }
      }
      /* 4.4c freezing of liquid, :1913-1928 */
      {
        const T frzmax = CLOUDSC_MX0((c.rthomo - t) * sc.zrldcp);
        if (CLOUDSC_RARE(frzmax > c.zepsec && qxfg[0] > c.zepsec)) {
          const T frz = mn(qxfg[0], frzmax);
          a01 = a01 + frz;
        }
      }
      /* 4.5 evaporation of rain, Abel & Boutle, :1984-2038 */
      /* zzrh0 is shared by the rain and snow branches, :1990-1991 == :2051-2052 */
      T zzrh0 = c.rprecrhmax + CLOUDSC_DIVR(sc.one_m_rprecrhmax * s.covpmax, mx(c.zepsec, one_m_a), r_1ma);
      zzrh0 = CLOUDSC_MN1(mx(zzrh0, c.rprecrhmax));
      {
        const T zzrh = mn(c.p08, zzrh0);
        bool rain_evap = covpclr > c.zepsec && qxfg[2] > c.zepsec;
        T qsliq = zero, qe = zero;
        if (CLOUDSC_RARE(rain_evap)) {
          const T foeeliqt = CLOUDSC_MNH(CLOUDSC_DIVR(c.r2es * e_liq, pap, r_pap));   /* :607-609 */
          const T cor_l = CLOUDSC_RCP(one - c.retv * foeeliqt);
          (void)cor_l;
          qsliq = CLOUDSC_DIVR(foeeliqt, one - c.retv * foeeliqt, cor_l);
          qe = CLOUDSC_MX0R(mn(qv, qsliq));
          rain_evap = qe < zzrh * qsliq;
        }
        if (CLOUDSC_RARE(rain_evap)) {
          const T preclr = CLOUDSC_DIV(qxfg[2], s.covptot);
          const T fallcorr = CLOUDSC_POW(CLOUDSC_DIVR(c.rdensref, rho, r_rho), c.p04);
          const T esatliq = sc.rv_rd * (c.r2es * e_liq);
          const T lambda = CLOUDSC_POW(CLOUDSC_DIV(c.rcl_fac1, rho * preclr), c.rcl_fac2);
          const T evap_denom = c.rcl_cdenom1 * esatliq - c.rcl_cdenom2 * t * esatliq + (c.rcl_cdenom3 * CLOUDSC_CUBE(t)) * pap;
          const T corr2 = CLOUDSC_DIV(CLOUDSC_POW15(CLOUDSC_DIVR(t, T(273.0), c.r_273)) * T(393.0), t + T(120.0));
          const T subsat = CLOUDSC_MX0(zzrh * qsliq - qe);
          const T beta = ((((CLOUDSC_DIV(T(0.5), qsliq) * (t * t)) * esatliq) * c.rcl_const1r) * CLOUDSC_DIV(corr2, evap_denom)) *
                         (CLOUDSC_DIV(c.p078, CLOUDSC_POW(lambda, c.rcl_const4r)) +
                          CLOUDSC_DIV(c.rcl_const2r * CLOUDSC_SQRT(rho * fallcorr), CLOUDSC_SQRT(corr2) * CLOUDSC_POW(lambda, c.rcl_const3r)));
          const T denom = one + beta * ptsphy;
          const T dpevap = CLOUDSC_DIV(((covpclr * beta) * ptsphy) * subsat, denom);
          const T evap = mn(dpevap, qxfg[2]);
          a24 = a24 + evap;
          s.covptot = mx(c.rcovpmin, s.covptot - CLOUDSC_MX0R(CLOUDSC_DIV((s.covptot - a) * evap, qxfg[2])));
          qxfg[2] = qxfg[2] - evap;
        }
      }
      /* 4.5 evaporation of snow, Sundqvist-type, :2050-2087 */
      {
        const T zzrh = zzrh0;
        T qe = CLOUDSC_DIVR(qv - a * qsice, mx(c.zepsec, one_m_a), r_1ma);
        qe = CLOUDSC_MX0R(mn(qe, qsice));
        if (covpclr > c.zepsec && qxfg[3] > c.zepsec && qe < zzrh * qsice) {
          const T cd = s.covptot * dtgdp;
          const T preclr = CLOUDSC_DIV(qxfg[3] * covpclr, cs(mx(ab(cd), c.zepsilon), cd));
          const T beta1 = CLOUDSC_DIV(CLOUDSC_DIVR(CLOUDSC_SQRT(CLOUDSC_DIVR(pap, paph_surf, r_paph_surf)), c.rvrfactor, c.r_rvrfactor) * preclr, mx(covpclr, c.zepsec));
          const T beta = sc.rg_rpecons * CLOUDSC_POW(beta1, c.p05777);
          const T denom = one + (beta * ptsphy) * corqsice;
          const T dpr = (CLOUDSC_DIV((covpclr * beta) * (qsice - qe), denom) * dp) * sc.zrg_r;
          const T dpevap = dpr * dtgdp;
          const T evap = mn(dpevap, qxfg[3]);
          a34 = a34 + evap;
          s.covptot = mx(c.rcovpmin, s.covptot - CLOUDSC_MX0R(CLOUDSC_DIV((s.covptot - a) * evap, qxfg[3])));
          qxfg[3] = qxfg[3] - evap;
        }
      }
      /* evaporate small precipitation amounts, :2144-2158 */
      if (qxfg[2] < c.rlmin) { a24 = a24 + qxfg[2]; }
      if (qxfg[3] < c.rlmin) { a34 = a34 + qxfg[3]; }

      /* ================= 5.1 solver for cloud cover, :2168-2180 =========================== */
      T anew = CLOUDSC_DIV(a + solac, one + solab);
      anew = CLOUDSC_MN1(anew);
      if (anew < c.ramin) anew = zero;
      T aorig_e = aorig;
      if constexpr (kReload) aorig_e = io.template ld<IN_PA>(k) + ptsphy * io.template ld<IN_TA>(k);
      const T da = anew - aorig_e;
      s.anewm1 = anew;

      /* ================= 5.2 sink limiter (:2233-2286) and 5.2.2 implicit solve (:2294-2412) ===== */
      /* (functions above: sink_limiter(), implicit_solve(); unit-tested against a dense restatement of the reference's
       * statements in tests/test_unit_processes.py) */
      T qxn[5];
      {
        T cs0, cs1, cs2, cs3, cs4;
        sink_limiter<T>(c, a01, a02, a03, a04, a12, a14, a23, a24, a34, d0, d1, d2, d3, qx, qv, cs0, cs1, cs2, cs3, cs4);
        implicit_solve<T>(c, qx, qv, cs0, cs1, cs2, cs3, cs4, qb00, qb11, qb03, qb13, fallsink1, fallsink2, fallsink3, qxn);
      }
      /* variables needed for next level, :2417-2423 */
      s.qlnm1 = qxn[0]; s.qinm1 = qxn[1];
#pragma unroll
      for (int m = 0; m < 4; ++m) qxn2d[m] = qxn[m];

      /* 5.3 precipitation/sedimentation fluxes to next level, :2430-2448 */
      pf_i_n = (fallsink1 * qxn[1]) * rdtgdp;
      pf_r_n = (fallsink2 * qxn[2]) * rdtgdp;
      pf_s_n = (fallsink3 * qxn[3]) * rdtgdp;
      if (pf_s_n + pf_r_n < c.zepsec) s.covptot = zero;
      park_st<PARK_COVPTOT>(io, s.covptot, s.covptot);
      park_st<PARK_COVPMAX>(io, s.covpmax, s.covpmax);

      /* ================= 6. tendencies, :2456-2506 ======================================== */
      {
        T fluxq = (psupsrce0 + convsrce0) - (zero + convsink) * qxn[0];
        tend_t = tend_t + (c.ralvdcp * (qxn[0] - qx[0] - fluxq)) * zqtmst;
        fluxq = ((psupsrce1 + convsrce1) + fallsrce1) - (fallsink1 + convsink) * qxn[1];
        tend_t = tend_t + (c.ralsdcp * (qxn[1] - qx[1] - fluxq)) * zqtmst;
        fluxq = fallsrce2 - fallsink2 * qxn[2];
        tend_t = tend_t + (c.ralvdcp * (qxn[2] - qx[2] - fluxq)) * zqtmst;
        fluxq = fallsrce3 - fallsink3 * qxn[3];
        tend_t = tend_t + (c.ralsdcp * (qxn[3] - qx[3] - fluxq)) * zqtmst;
      }
      {
        T q0[4] = {qx0[0], qx0[1], qx0[2], qx0[3]};
        if constexpr (kReload) {
          q0[0] = io.template ld<IN_PCLV0>(k) + ptsphy * io.template ld<IN_TCLD0>(k);
          q0[1] = io.template ld<IN_PCLV1>(k) + ptsphy * io.template ld<IN_TCLD1>(k);
          q0[2] = io.template ld<IN_PCLV2>(k) + ptsphy * io.template ld<IN_TCLD2>(k);
          q0[3] = io.template ld<IN_PCLV3>(k) + ptsphy * io.template ld<IN_TCLD3>(k);
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) tcld[m] = (qxn[m] - q0[m]) * zqtmst;
      }
      tend_q = tend_q + (qxn[4] - qv) * zqtmst;
      tend_a = da * zqtmst;
      covptot_out = s.covptot;
    } /* k >= ktop */

    /* ---- write the level's tendencies (single store per element) */
    io.template st<OUT_TT>(k, tend_t);
    io.template st<OUT_TQ>(k, tend_q);
    io.template st<OUT_TA>(k, tend_a);
    io.template st<OUT_TCLD0>(k, tcld[0]); io.template st<OUT_TCLD1>(k, tcld[1]);
    io.template st<OUT_TCLD2>(k, tcld[2]); io.template st<OUT_TCLD3>(k, tcld[3]);
    io.template st<OUT_TCLD4>(k, zero);    /* cloudsc_driver.c:200 */
    io.template st<OUT_COVPTOT>(k, covptot_out); /* :2505; zero above ncldtop per cloudsc_driver.c:199 */
    io.template st<OUT_PLUDE>(k, plude_out);     /* in/out: the loaded value where :970-980 does not apply */

    /* ================= 8. flux diagnostics for row k+1, :2521-2582 ========================= */
    {
      const T pfplsl = pf_r_n + zero;     /* zpfplsx[r] + zpfplsx[l] */
      const T pfplsn = pf_s_n + pf_i_n;   /* zpfplsx[s] + zpfplsx[i] */
      io.template st<OUT_FPLSL>(k + 1, pfplsl);
      io.template st<OUT_FPLSN>(k + 1, pfplsn);
      io.template st<OUT_FHPSL>(k + 1, -c.rlvtt * pfplsl);
      io.template st<OUT_FHPSN>(k + 1, -c.rlstt * pfplsn);

      const T gdph_r = (sc.neg_zrg_r * (paph_n - s.paph_k)) * zqtmst;
      const T pvfl = io.template ld<IN_PVFL>(k), pvfi = io.template ld<IN_PVFI>(k);
      /* rain/snow rows copy the liquid/ice running sums of the level above (:2550-2555).
       * zlneg[m] (:526,550: the condensate removed by the tidy-up, else 0) == qx0[m] - qx[m] exactly. */
      T q0[4] = {qx0[0], qx0[1], qx0[2], qx0[3]};
      if constexpr (kReload) {
        q0[0] = io.template ld<IN_PCLV0>(k) + ptsphy * io.template ld<IN_TCLD0>(k);
        q0[1] = io.template ld<IN_PCLV1>(k) + ptsphy * io.template ld<IN_TCLD1>(k);
        q0[2] = io.template ld<IN_PCLV2>(k) + ptsphy * io.template ld<IN_TCLD2>(k);
        q0[3] = io.template ld<IN_PCLV3>(k) + ptsphy * io.template ld<IN_TCLD3>(k);
      }
      const T o_fsqlf = park_ld<PARK_FSQLF>(io, s.fsqlf), o_fsqif = park_ld<PARK_FSQIF>(io, s.fsqif);
      const T o_fcqlng = park_ld<PARK_FCQLNG>(io, s.fcqlng), o_fcqnng = park_ld<PARK_FCQNNG>(io, s.fcqnng);
      const T o_fsqltur = park_ld<PARK_FSQLTUR>(io, s.fsqltur), o_fsqitur = park_ld<PARK_FSQITUR>(io, s.fsqitur);
      const T fsqrf = o_fsqlf + (qxn2d[2] - q0[2]) * gdph_r;
      const T fcqrng = o_fcqlng + (q0[2] - qx[2]) * gdph_r;
      const T fsqsf = o_fsqif + (qxn2d[3] - q0[3]) * gdph_r;
      const T fcqsng = o_fcqnng + (q0[3] - qx[3]) * gdph_r;
      const T n_fsqlf = o_fsqlf + (qxn2d[0] - q0[0] + pvfl * ptsphy - alfa_w * plude_out) * gdph_r;
      const T n_fcqlng = o_fcqlng + (q0[0] - qx[0]) * gdph_r;
      const T n_fsqltur = o_fsqltur + (pvfl * ptsphy) * gdph_r;
      const T n_fsqif = o_fsqif + (qxn2d[1] - q0[1] + pvfi * ptsphy - (one - alfa_w) * plude_out) * gdph_r;
      const T n_fcqnng = o_fcqnng + (q0[1] - qx[1]) * gdph_r;
      const T n_fsqitur = o_fsqitur + (pvfi * ptsphy) * gdph_r;
      park_st<PARK_FSQLF>(io, s.fsqlf, n_fsqlf); park_st<PARK_FSQIF>(io, s.fsqif, n_fsqif);
      park_st<PARK_FCQLNG>(io, s.fcqlng, n_fcqlng); park_st<PARK_FCQNNG>(io, s.fcqnng, n_fcqnng);
      park_st<PARK_FSQLTUR>(io, s.fsqltur, n_fsqltur); park_st<PARK_FSQITUR>(io, s.fsqitur, n_fsqitur);
      io.template st<OUT_FSQLF>(k + 1, n_fsqlf);   io.template st<OUT_FSQIF>(k + 1, n_fsqif);
      io.template st<OUT_FSQRF>(k + 1, fsqrf);     io.template st<OUT_FSQSF>(k + 1, fsqsf);
      io.template st<OUT_FCQLNG>(k + 1, n_fcqlng); io.template st<OUT_FCQNNG>(k + 1, n_fcqnng);
      io.template st<OUT_FCQRNG>(k + 1, fcqrng);   io.template st<OUT_FCQSNG>(k + 1, fcqsng);
      io.template st<OUT_FSQLTUR>(k + 1, n_fsqltur); io.template st<OUT_FSQITUR>(k + 1, n_fsqitur);
    }

    /* ---- shift the level-(k-1) cache */
    s.tp1_m1 = t; s.a_m1 = a; s.pap_m1 = pap; s.paph_k = paph_n; s.mf_k = mf_n;
    park_st<PARK_PF_I>(io, s.pf_i, pf_i_n); park_st<PARK_PF_R>(io, s.pf_r, pf_r_n); park_st<PARK_PF_S>(io, s.pf_s, pf_s_n);
    io.level_end(k);
  }
  if constexpr (io_ranged<IO>::value) {
    if (k_end < klev) { io.save_carry(s); return; }          /* cut: the rest of the column runs later */
  }
  io.st_rainfrac(park_ld<PARK_RAINFRAC>(io, s.rainfrac));
}
#undef paph_surf
#undef r_paph_surf
#undef kk_num_pow
#undef lcrit_kk

} /* namespace cloudsc */
#endif
