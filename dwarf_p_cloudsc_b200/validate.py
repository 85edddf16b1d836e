"""The dwarf's error table (``cloudsc_validate``) with an explicit pass/fail bar.

Semantics follow src/cloudsc_c/cloudsc/cloudsc_validate.c:20-115 per field over the VALID columns:
min, max, AbsMaxErr = max|f-ref|, AvgAbsErr/GP = sum|f-ref| / ngptot, and the "MaxRelErr-%" column
    0                       if sum|f-ref| < eps
    sum|err| / (1 + zsum)   if zsum < eps
    sum|err| / zsum         otherwise                     (x100; " !!!!" if > 10 eps)
where the reference accumulates ``zsum += abs(ref)`` with the *integer* ``abs`` (cloudsc_validate.c:74,109,147;
SURVEY App. C.1), i.e. every |ref| < 1 counts as 0.  We print that reference-compatible number AND the
true relative L1 error, and judge on the latter:

    FP64 bar:   sum|err|/sum|ref| <= 1e-10   and   AbsMaxErr <= 1e-9 * max|ref|
    FP32 (the reference never validates its SP builds, .github/workflows/build.yml:172, so the bars are stated here,
    from measurement, profiles/r2_fp32_three_way.txt): single precision moves threshold decisions of the scheme
    (tests like `zqx < rlmin`, `a < ramin`, `zsupsat > zepsec` on values that sit within a few float ulps of the
    threshold), so whole condensation events appear or vanish at isolated points and the flux sums integrate them down
    the column.  That is a property of CLOUDSC in FP32, not of this kernel: the reference's own kernel re-typed to
    float (oracle/_ref32) differs from its FP64 self by up to 1.2e-1 (PFSQLF).  Two bars therefore:
        "fp32ref"  GPU-FP32 against the CPU-FP32 reference: FP32_L1_BARS_VS_REF32 (measured worst 5e-6 .. 5.4e-2)
        "fp32"     any FP32 result against the FP64 reference: FP32_L1_BARS (measured worst 5e-6 .. 1.2e-1, for the
                   device path and for the CPU-FP32 reference alike)
"""
from __future__ import annotations

from typing import Dict, List, NamedTuple

import numpy as np

from .state import VALIDATE_ORDER, unblock

EPS = float(np.finfo(np.float64).eps)
FP64_L1_TOL = 1.0e-10
FP64_MAX_TOL = 1.0e-9
FP32_L1_TOL = 1.0e-2          # default per-field bar of the FP32 build
FP32_L1_BARS = {"PFSQLF": 3e-1, "PFSQRF": 3e-1, "PFSQIF": 1e-1, "PFSQSF": 1e-1, "PFCQLNG": 3e-1, "PFCQNNG": 3e-1,
                "PFCQRNG": 3e-1, "PFCQSNG": 3e-1, "PFPLSN": 2e-2, "PFHPSN": 2e-2, "PFPLSL": 8e-2, "PFHPSL": 8e-2,
                "TENDENCY_LOC_Q": 5e-2, "TENDENCY_LOC_T": 5e-2, "TENDENCY_LOC_CLD": 5e-2, "TENDENCY_LOC_A": 1e-2,
                "PCOVPTOT": 3e-2, "PLUDE": 1e-4}   # (PCOVPTOT, PFPLSL/PFHPSL: set by the synthetic warm-phase input sets)

# GPU-FP32 vs the CPU-FP32 reference (oracle/_ref32): ~3x the worst value measured over the shipped data and three
# synthetic input sets (profiles/r2_fp32_three_way.txt)
FP32_L1_BARS_VS_REF32 = {"PLUDE": 1e-6, "PCOVPTOT": 1.5e-2, "PRAINFRAC_TOPRFZ": 1e-6, "PFSQLF": 2e-2, "PFSQRF": 2e-2,
                         "PFSQIF": 1.2e-2, "PFSQSF": 1.2e-2, "PFCQLNG": 1.5e-1, "PFCQNNG": 1.5e-1, "PFCQRNG": 1.5e-1,
                         "PFCQSNG": 1.5e-1, "PFSQLTUR": 1e-6, "PFSQITUR": 1e-6, "PFPLSL": 5e-2, "PFHPSL": 5e-2,
                         "PFPLSN": 1e-2, "PFHPSN": 1e-2, "TENDENCY_LOC_A": 5e-5, "TENDENCY_LOC_Q": 3e-2,
                         "TENDENCY_LOC_T": 2e-2, "TENDENCY_LOC_CLD": 1.5e-2}

_PRINT_NAME = {"TENDENCY_LOC_A": "TENDENCY_LOC%A", "TENDENCY_LOC_Q": "TENDENCY_LOC%Q",
               "TENDENCY_LOC_T": "TENDENCY_LOC%T", "TENDENCY_LOC_CLD": "TENDENCY_LOC%CLD"}


class FieldError(NamedTuple):
    name: str
    minval: float
    maxval: float
    maxerr: float
    errsum: float
    refsum_true: float      # sum |ref|
    refsum_c: float         # sum (int)|ref|, the C validator's zsum
    refmax: float
    avgpgp: float
    relerr_c: float         # the reference's "MaxRelErr-%" (already x100)
    iopt: int
    rel_l1: float           # true sum|err|/sum|ref|
    flagged: bool           # the reference's " !!!!"

    def passes(self, precision: str = "fp64") -> bool:
        if not np.isfinite([self.maxerr, self.errsum, self.minval, self.maxval]).all():
            return False
        if precision == "fp32":
            return self.rel_l1 <= FP32_L1_BARS.get(self.name, FP32_L1_TOL)
        if precision == "fp32ref":
            return self.rel_l1 <= FP32_L1_BARS_VS_REF32.get(self.name, FP32_L1_TOL)
        return self.rel_l1 <= FP64_L1_TOL and self.maxerr <= FP64_MAX_TOL * max(self.refmax, 1e-300)


def field_error(name: str, got: np.ndarray, ref: np.ndarray, ngptot: int) -> FieldError:
    """got/ref: arrays over valid columns only, any shape (same for both)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    diff = np.abs(got - ref)
    errsum = float(diff.sum())
    refsum_true = float(np.abs(ref).sum())
    refsum_c = float(np.trunc(np.abs(ref)).sum())     # int abs() on a double truncates toward zero
    if errsum < EPS:
        rel, iopt = 0.0, 1
    elif refsum_c < EPS:
        rel, iopt = errsum / (1.0 + refsum_c), 2
    else:
        rel, iopt = errsum / refsum_c, 3
    rel_l1 = 0.0 if errsum == 0.0 else (errsum / refsum_true if refsum_true > 0 else float("inf"))
    return FieldError(name, float(got.min()), float(got.max()), float(diff.max()), errsum, refsum_true, refsum_c,
                      float(np.abs(ref).max()), errsum / ngptot, 100.0 * rel, iopt, rel_l1, rel > 10.0 * EPS)


def _merge(name: str, parts, ngptot: int) -> FieldError:
    """Combine partial sums (min, max, maxerr, errsum, refsum_true, refsum_c, refmax) into one FieldError."""
    mn = min(p[0] for p in parts); mx = max(p[1] for p in parts); maxerr = max(p[2] for p in parts)
    errsum = float(sum(p[3] for p in parts)); refsum_true = float(sum(p[4] for p in parts))
    refsum_c = float(sum(p[5] for p in parts)); refmax = max(p[6] for p in parts)
    if errsum < EPS:
        rel, iopt = 0.0, 1
    elif refsum_c < EPS:
        rel, iopt = errsum / (1.0 + refsum_c), 2
    else:
        rel, iopt = errsum / refsum_c, 3
    rel_l1 = 0.0 if errsum == 0.0 else (errsum / refsum_true if refsum_true > 0 else float("inf"))
    return FieldError(name, float(mn), float(mx), float(maxerr), errsum, refsum_true, refsum_c, float(refmax),
                      errsum / ngptot, 100.0 * rel, iopt, rel_l1, rel > 10.0 * EPS)


def _partial(got: np.ndarray, ref: np.ndarray):
    diff = np.abs(got - ref)
    aref = np.abs(ref)
    # sums over a broadcast reference count every replica: broadcast explicitly where shapes differ
    if aref.shape != diff.shape:
        reps = diff.size // aref.size
        return (got.min(), got.max(), diff.max(), diff.sum(dtype=np.float64), aref.sum(dtype=np.float64) * reps,
                np.trunc(aref).sum(dtype=np.float64) * reps, aref.max())
    return (got.min(), got.max(), diff.max(), diff.sum(dtype=np.float64), aref.sum(dtype=np.float64),
            np.trunc(aref).sum(dtype=np.float64), aref.max())


def validate_replicated(fields: Dict[str, np.ndarray], reference_columns: Dict[str, np.ndarray], ngptot: int, nproma: int,
                        first_column: int = 0) -> List[FieldError]:
    """The dwarf's check at full size: EVERY column of blocked output arrays [nb, ..., nproma] against the reference
    data set replicated the way load_reference / expand_* replicate it (global column g -> g mod nlon,
    load_state.c:123-138,701-760; `first_column` = this rank's offset, expand_mod.F90:30-46).  The replicated reference
    repeats every lcm(nproma, nlon) columns, so it is built for one period of blocks and broadcast -- no full-size copy."""
    from .state import expand
    out = []
    nb = (ngptot + nproma - 1) // nproma
    for name in VALIDATE_ORDER:
        got = np.asarray(fields[name])
        if got.dtype != np.float64:
            got = got.astype(np.float64)
        ref = np.asarray(reference_columns[name], dtype=np.float64)
        nlon = ref.shape[-1]
        period = int(np.lcm(nproma, nlon)) // nproma                   # blocks after which the pattern repeats
        exp = expand(ref, min(nb, period) * nproma, nproma, first_column)     # [min(nb, period), ..., nproma]
        valid_last = ngptot - (nb - 1) * nproma
        nfullblk = nb if valid_last == nproma else nb - 1               # blocks whose columns are all valid
        parts = []
        whole = (nfullblk // period) * period if period <= nfullblk else 0
        if whole:
            g = got[:whole].reshape((whole // period, period) + got.shape[1:])
            parts.append(_partial(g, exp[None]))
        for b in range(whole, nfullblk):
            parts.append(_partial(got[b], exp[b % period]))
        if nfullblk < nb:
            parts.append(_partial(got[nb - 1][..., :valid_last], exp[(nb - 1) % period][..., :valid_last]))
        out.append(_merge(name, parts, ngptot))
    return out


def validate(fields: Dict[str, np.ndarray], reference: Dict[str, np.ndarray], ngptot: int) -> List[FieldError]:
    """fields/reference: blocked arrays [nb, ..., nproma] of the 21 validated outputs."""
    out = []
    for name in VALIDATE_ORDER:
        out.append(field_error(name, unblock(np.asarray(fields[name]), ngptot),
                               unblock(np.asarray(reference[name]), ngptot), ngptot))
    return out


def format_table(errs: List[FieldError], precision: str = "fp64") -> str:
    """The reference's table layout (cloudsc_validate.c:44,193) plus two columns of ours."""
    lines = [" %20s %s %20s %20s %20s %20s %20s   %12s %s" % (
        "Variable", "Dim", "MinValue", "MaxValue", "AbsMaxErr", "AvgAbsErr/GP", "MaxRelErr-%", "L1rel(true)", "ok")]
    for e in errs:
        lines.append(" %20s %dD%d %20.13e %20.13e %20.13e %20.13e %20.13e %s %12.5e %s" % (
            _PRINT_NAME.get(e.name, e.name), 2, e.iopt, e.minval, e.maxval, e.maxerr, e.avgpgp, e.relerr_c,
            " !!!!" if e.flagged else "     ", e.rel_l1, "PASS" if e.passes(precision) else "FAIL"))
    return "\n".join(lines)


def all_pass(errs: List[FieldError], precision: str = "fp64") -> bool:
    return all(e.passes(precision) for e in errs)
