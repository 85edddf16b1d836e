"""Per-process unit oracles: the three pieces of the level body whose device implementation is organised
differently from the reference -- the sink limiter on the antisymmetric ZSOLQA, the closed-form 5x5 implicit solve and the
Newton step of the saturation adjustment -- each called ON ITS OWN (host build of the device code, tests/host_emul) against a
dense, statement-by-statement restatement of the reference lines (oracle/unit_oracles.c: cloudsc_c.c:2233-2286, :2294-2412,
:1162-1168), on adversarial inputs the 100-column data set never produces (every species over-drawn at once, zero content,
huge transfer coefficients, saturated air).  Strict arithmetic: bit-identical.  Fast arithmetic (what the GPU runs): 1e-13."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_d = ctypes.POINTER(ctypes.c_double)
PAIRS = [(0, 1), (0, 2), (0, 3), (0, 4), (1, 2), (1, 4), (2, 3), (2, 4), (3, 4)]   # a01 a02 a03 a04 a12 a14 a23 a24 a34


@pytest.fixture(scope="module")
def libs(built):
    unit = ctypes.CDLL(os.path.join(ROOT, "oracle", "libcloudsc_unit.so"))
    strict = ctypes.CDLL(os.path.join(ROOT, "tests", "host_emul", "libcloudsc_emul.so"))
    fast = ctypes.CDLL(os.path.join(ROOT, "tests", "host_emul", "libcloudsc_emul_fast.so"))
    return unit, strict, fast


def _p(a):
    return a.ctypes.data_as(_d)


def _dense(a13):
    """the 13 stored values -> dense zsolqa[from][to] (antisymmetric pairs, diagonal sources)"""
    z = np.zeros((5, 5))
    for v, (x, y) in zip(a13[:9], PAIRS):
        z[x, y] = v
        z[y, x] = -v
    for m in range(4):
        z[m, m] = a13[9 + m]
    return z


def _draw(rng, case):
    """sources (13 values) and contents (5 species) of one grid point"""
    mag = 10.0 ** rng.uniform(-12, -3, 13)
    a = mag * rng.choice([-1.0, 1.0], 13)
    a[9:] = mag[9:] * rng.choice([-1.0, 1.0, 1.0, 0.0], 4)            # diagonal: mostly sources, some sinks, some nothing
    a[rng.random(13) < 0.3] = 0.0
    qx = 10.0 ** rng.uniform(-9, -3, 5)
    qx[4] = 10.0 ** rng.uniform(-6, -2)
    if case == "overdrawn":        # every species loses far more than it holds
        a[:9] = np.abs(a[:9]) * np.array([1, 1, 1, 1, 1, 1, 1, 1, 1]) * 1e3
        qx[:4] *= 1e-3
    elif case == "empty":          # nothing there at all: the limiter's max(., zepsec) floor decides
        qx[:4] = 0.0
    elif case == "quiet":          # sinks well inside the contents: nothing may be scaled
        a *= 1e-6
    return a, qx


@pytest.mark.parametrize("case", ["mixed", "overdrawn", "empty", "quiet"])
def test_sink_limiter_against_the_dense_restatement(libs, case):
    unit, strict, fast = libs
    rng = np.random.default_rng({"mixed": 1, "overdrawn": 2, "empty": 3, "quiet": 4}[case])
    scaled = 0
    for _ in range(400):
        a, qx = _draw(rng, case)
        z = _dense(a)
        z0 = z.copy()
        unit.unit_sink_limiter(_p(z), _p(qx), ctypes.c_double(1.0e-14))
        scaled += int(not np.array_equal(z, z0))
        assert np.array_equal(z, -z.T + 2 * np.diag(np.diag(z))), "the reference keeps ZSOLQA antisymmetric"
        want = np.array([z[x, y] for x, y in PAIRS] + [z[m, m] for m in range(4)])
        want_cs = np.array([sum_in_order(z[:, m]) for m in range(5)])
        for lib, exact in ((strict, True), (fast, False)):
            got, cs = a.copy(), np.zeros(5)
            lib.cloudsc_emul_unit_limiter(_p(got), _p(qx), _p(cs))
            if exact:
                assert np.array_equal(got, want), (case, got - want)
                assert np.array_equal(cs, want_cs), (case, cs - want_cs)
            else:
                # (the column sums cancel: their error is relative to the largest term, not to the sum)
                assert np.allclose(got, want, rtol=1e-13, atol=0.0) and np.allclose(cs, want_cs, rtol=0.0, atol=1e-13 * np.abs(z).max())
    assert (scaled == 0) == (case == "quiet"), (case, scaled)


def sum_in_order(col):
    s = 0.0
    for v in col:
        s = s + v
    return s


@pytest.mark.parametrize("case", ["mixed", "overdrawn", "empty"])
def test_implicit_solve_against_the_dense_lu(libs, case):
    """dense LHS + generic non-pivoting LU + back-substitution + clipping (cloudsc_c.c:2294-2412) vs the closed form for
    the four ZSOLQB entries the hard-coded process set can produce ((l,l) (i,i) (l,s) (i,s))"""
    unit, strict, fast = libs
    rng = np.random.default_rng({"mixed": 11, "overdrawn": 12, "empty": 13}[case])
    clipped = 0
    for _ in range(400):
        a, qx = _draw(rng, case)
        z = _dense(a)
        unit.unit_sink_limiter(_p(z), _p(qx), ctypes.c_double(1.0e-14))
        qb = 10.0 ** rng.uniform(-6, 3, 4) * (rng.random(4) < 0.7)         # qb00 qb11 qb03 qb13: up to huge (implicit sinks)
        fs = 10.0 ** rng.uniform(-4, 2, 3)
        zsolqb = np.zeros((5, 5))
        zsolqb[0, 0], zsolqb[1, 1], zsolqb[0, 3], zsolqb[1, 3] = qb
        zfallsink = np.array([0.0, fs[0], fs[1], fs[2], 0.0])
        want = np.zeros(5)
        unit.unit_implicit_solve(_p(z), _p(zsolqb), _p(qx), _p(zfallsink), ctypes.c_double(1.0e-14), _p(want))
        clipped += int((want[:4] == 0.0).any())
        cs = np.array([sum_in_order(z[:, m]) for m in range(5)])
        for lib, exact in ((strict, True), (fast, False)):
            got = np.zeros(5)
            lib.cloudsc_emul_unit_solve(_p(qx), _p(cs), _p(qb), _p(fs), _p(got))
            if exact:
                assert np.array_equal(got, want), (case, got, want)
            else:
                assert np.allclose(got, want, rtol=1e-13, atol=1e-27), (case, got, want)
        assert np.isfinite(want).all() and (want[:4] >= 0.0).all()
    assert clipped > 0


def test_saturation_adjustment_step_against_the_inlined_statement_functions(libs, input_data):
    """one Newton step, cloudsc_c.c:1162-1168, over the whole (T, p, q) range of the troposphere incl. the mixed-phase
    interval, the homogeneous-freezing end and air at the 0.5 cap of the saturation mixing ratio"""
    unit, strict, fast = libs
    from dwarf_p_cloudsc_b200 import params
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    kk = np.array([k.rtt, k.r2es, k.r3les, k.r3ies, k.r4les, k.r4ies, k.r5alvcp, k.r5alscp, k.ralvdcp, k.ralsdcp, k.rtwat, k.rtice,
                   k.rtwat_rtice_r, k.retv])
    rng = np.random.default_rng(21)
    for lib in (strict, fast):
        lib.cloudsc_emul_unit_satadj.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, _d, _d]
    worst = 0.0
    for _ in range(3000):
        t = rng.uniform(160.0, 320.0)
        pap = 10.0 ** rng.uniform(1.0, 5.02)                      # 10 Pa .. 1050 hPa: the top levels hit the 0.5 cap
        qsm = 10.0 ** rng.uniform(-7, -1.5)
        qp = 1.0 / pap
        wt, wq = ctypes.c_double(t), ctypes.c_double(qsm)
        unit.unit_satadj_step(_p(kk), ctypes.c_double(qp), ctypes.byref(wt), ctypes.byref(wq))
        for lib, exact in ((strict, True), (fast, False)):
            gt, gq = ctypes.c_double(t), ctypes.c_double(qsm)
            assert lib.cloudsc_emul_unit_satadj(ctypes.addressof(p), ctypes.addressof(k), qp, ctypes.byref(gt), ctypes.byref(gq)) == 0
            if exact:
                assert gt.value == wt.value and gq.value == wq.value, (t, pap, qsm)
            else:
                worst = max(worst, abs(gt.value - wt.value) / abs(wt.value), abs(gq.value - wq.value) / max(abs(wq.value), abs(qsm)))
    assert worst < 1e-13, worst
