"""CPU-side parity chain (`-m "not gpu"`):

    reference.h5 (Fortran kernel)  <-3e-13->  compiled reference C kernel (oracle/_ref, when present)
         <-bit exact->  C restatement (oracle/cloudsc_oracle.c)
         <-bit exact->  host build of the DEVICE column code (tests/host_emul, FP contraction off)

so that on the GPU only FMA contraction and the CUDA libm separate the kernel from the reference.
"""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = ["", "warm", "hot+land", "conv+land", "rain+warm", "moist+tend",
            "warm+land+conv+rain+moist+tend", "hot+rain+moist+conv", "sparse", "sparse+warm+conv"]


def _oracles(built):
    import pyoracle
    out = []
    for cls in (pyoracle.RefOracle, pyoracle.PortOracle):
        try:
            out.append(cls())
        except (FileNotFoundError, OSError):
            pass
    assert out, "no CPU oracle could be loaded"
    return out


def test_oracles_reproduce_reference_h5(built, input_data):
    """Pins the oracle: every CPU checker vs the golden vectors of the reference's own data set.
    Bars = what the compiled reference kernel itself achieves against the Fortran-generated file
    (SURVEY App. D): L1-relative <= 1e-15, point-wise relative <= 1e-12 where |ref| > 1e-3 max|ref|."""
    from dwarf_p_cloudsc_b200 import params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    ref = state.load_reference()
    for orc in _oracles(built):
        for nproma in (100, 16):
            F = state.blocked_fields(cols, 100, nproma, klev)
            orc.run(F, p, k, 100, nproma, klev, scalars["PTSPHY"], 2)
            for name in state.VALIDATE_ORDER:
                got, exp = state.unblock(F[name], 100), ref[name]
                err = np.abs(got - exp)
                denom = np.abs(exp).sum()
                assert err.sum() <= 1e-15 * denom + 1e-300, (orc.kind, name, err.sum() / max(denom, 1e-300))
                big = np.abs(exp) > 1e-3 * np.abs(exp).max()
                if big.any():
                    assert (err[big] / np.abs(exp[big])).max() <= 1e-12, (orc.kind, name)


def test_port_is_bit_identical_to_reference_build(built, input_data):
    import pyoracle
    try:
        ref, port = pyoracle.RefOracle(), pyoracle.PortOracle()
    except (FileNotFoundError, OSError):
        pytest.skip("needs both oracle/_ref and the port (reference checkout not present)")
    from dwarf_p_cloudsc_b200 import params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    for variant in VARIANTS:
        c2 = state.perturb(cols, variant)
        A = state.blocked_fields(c2, 150, 32, klev)
        B = {n: a.copy() for n, a in A.items()}
        ref.run(A, p, k, 150, 32, klev, scalars["PTSPHY"], 1)
        port.run(B, p, k, 150, 32, klev, scalars["PTSPHY"], 1)
        for f in params.OUTPUT_FIELDS:
            a, b = state.unblock(A[f.name], 150), state.unblock(B[f.name], 150)
            assert np.isfinite(a).all()
            assert np.array_equal(a, b), (variant, f.name, np.abs(a - b).max())


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("nssopt", [1, 0, 2, 3])
def test_device_column_code_host_build_is_bit_identical(built, oracle, input_data, variant, nssopt):
    """The k-cached / sparse-solve re-organisation is value-identical to the reference algorithm."""
    if nssopt != 1 and variant not in ("moist+tend", "warm+land+conv+rain+moist+tend", "sparse"):
        pytest.skip("nssopt sweep only on the variants that reach 3.4b(2)")
    import pyemul
    from dwarf_p_cloudsc_b200 import params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    p.nssopt = nssopt
    c2 = state.perturb(cols, variant)
    ngptot, nproma = 100, 16       # 7 blocks, ragged last block (4 valid columns)
    A = state.blocked_fields(c2, ngptot, nproma, klev)
    B = {n: a.copy() for n, a in A.items()}
    oracle.run(A, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 1)
    pyemul.Emul().run(B, p, k, ngptot, nproma, klev, scalars["PTSPHY"])
    for f in params.OUTPUT_FIELDS:
        a, b = state.unblock(A[f.name], ngptot), state.unblock(B[f.name], ngptot)
        assert np.array_equal(a, b), (variant, f.name, float(np.abs(a - b).max()))


@pytest.mark.parametrize("variant", ["", "warm+land+conv+rain+moist+tend", "hot+rain+moist+conv", "sparse"])
@pytest.mark.parametrize("fast", [False, True])
def test_parked_state_code_path_is_value_identical(built, oracle, input_data, variant, fast):
    """The TMA kernels keep long-lived sweep state (flux sums, precipitation carry, per-column constants) in per-thread
    scratch of the IO policy instead of registers.  Host build of that code path (scratch poisoned before every column):
    strict mode bit-identical to the reference, fast mode bit-identical to the register path of the same build."""
    import pyemul
    from dwarf_p_cloudsc_b200 import params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    c2 = state.perturb(cols, variant)
    ngptot, nproma = 100, 16
    A = state.blocked_fields(c2, ngptot, nproma, klev)
    B = {n: a.copy() for n, a in A.items()}
    em = pyemul.Emul(fast=fast)
    if fast:
        em.run(A, p, k, ngptot, nproma, klev, scalars["PTSPHY"])
    else:
        oracle.run(A, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 1)
    em.run(B, p, k, ngptot, nproma, klev, scalars["PTSPHY"], parked=True)
    for f in params.OUTPUT_FIELDS:
        a, b = state.unblock(A[f.name], ngptot), state.unblock(B[f.name], ngptot)
        assert np.isfinite(b).all(), f.name
        assert np.array_equal(a, b), (variant, f.name, float(np.abs(a - b).max()))


@pytest.mark.parametrize("seed", [1, 2, 3, 5, 8, 13, 21, 34])
def test_seed_sweep_of_the_perturbed_inputs(built, oracle, input_data, seed):
    """Eight more draws of the two busiest synthetic input sets: reference build == plain-C port == host build of the
    device column code (register path and parked path), bit for bit."""
    import pyemul
    import pyoracle
    from dwarf_p_cloudsc_b200 import params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    ngptot, nproma = 100, 32
    em = pyemul.Emul()
    port = pyoracle.PortOracle()
    for variant in ("warm+land+conv+rain+moist+tend", "sparse+hot+rain+tend"):
        c2 = state.perturb(cols, variant, seed=seed)
        A = state.blocked_fields(c2, ngptot, nproma, klev)
        runs = []
        for how in ("ref", "port", "emul", "parked"):
            B = {n: a.copy() for n, a in A.items()}
            if how == "ref":
                oracle.run(B, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 1)
            elif how == "port":
                port.run(B, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 1)
            else:
                em.run(B, p, k, ngptot, nproma, klev, scalars["PTSPHY"], parked=(how == "parked"))
            runs.append(B)
        for f in params.OUTPUT_FIELDS:
            ref = state.unblock(runs[0][f.name], ngptot)
            assert np.isfinite(ref).all(), (variant, f.name)
            for how, B in zip(("port", "emul", "parked"), runs[1:]):
                assert np.array_equal(ref, state.unblock(B[f.name], ngptot)), (variant, seed, how, f.name)


@pytest.mark.parametrize("variant", VARIANTS)
def test_fast_arithmetic_mode_stays_inside_the_fp64_tolerance(built, oracle, input_data, variant):
    """CLOUDSC_STRICT=0 is the arithmetic of the product kernels: shared reciprocals, table-driven exp, atanh-series log,
    pow = exp(y log x), one-reciprocal Newton step, lazy evaluation of what is not read.  Its host build (libm division
    and sqrt in place of the MUFU seeds + Newton steps, no FMA contraction) must stay inside the per-field FP64
    tolerance of the GPU parity tests on every synthetic input set, three draws each."""
    import pyemul
    from dwarf_p_cloudsc_b200 import params, state, validate
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    ngptot, nproma = 100, 32
    em = pyemul.Emul(fast=True)
    for seed in (1234, 7, 99):
        c2 = state.perturb(cols, variant, seed=seed)
        R = state.blocked_fields(c2, ngptot, nproma, klev)
        G = {n: a.copy() for n, a in R.items()}
        oracle.run(R, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 1)
        em.run(G, p, k, ngptot, nproma, klev, scalars["PTSPHY"], parked=True)
        errs = validate.validate(G, R, ngptot)
        assert validate.all_pass(errs), (variant, seed, "\n" + validate.format_table(errs))
        assert max(e.rel_l1 for e in errs) < 1e-12, (variant, seed, max(e.rel_l1 for e in errs))   # measured <= 3.1e-14


def test_perturbations_reach_the_cold_branches(built, oracle, input_data):
    """The synthetic variants must actually exercise what the shipped data leaves cold."""
    from dwarf_p_cloudsc_b200 import params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    c2 = state.perturb(cols, "warm+land+conv+rain+moist+tend")
    F = state.blocked_fields(c2, 100, 100, klev)
    oracle.run(F, p, k, 100, 100, klev, scalars["PTSPHY"], 1)
    assert c2["PT"].max() > 280.0
    assert F["PFPLSL"].max() > 0.0                 # rain reaches the ground: melting + warm rain
    assert np.count_nonzero(F["PRAINFRAC_TOPRFZ"]) > 0   # rain freezing layer found
    assert F["PFHPSL"].min() < 0.0
    base = state.blocked_fields(cols, 100, 100, klev)
    oracle.run(base, p, k, 100, 100, klev, scalars["PTSPHY"], 1)
    assert base["PFPLSL"].max() == 0.0 and np.count_nonzero(base["PRAINFRAC_TOPRFZ"]) == 0


def test_float32_host_build_tracks_float64(built, input_data):
    """FP32 instantiation of the same column code (config C5).  Single precision moves threshold decisions of the
    scheme (values a few float ulps from rlmin / ramin / zepsec tests), so against FP64 the bar is only "finite and the
    same climate": measured L1-relative 5e-6 .. 1.2e-1 per field -- the same distance the reference's own kernel shows
    when re-typed to float (test_fp32_reference_pins_the_single_precision_column_code)."""
    import pyemul
    from dwarf_p_cloudsc_b200 import params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    A = state.blocked_fields(cols, 100, 100, klev)
    B = state.blocked_fields(cols, 100, 100, klev, dtype=np.float32)
    em = pyemul.Emul()
    em.run(A, p, k, 100, 100, klev, scalars["PTSPHY"])
    em.run(B, p, k, 100, 100, klev, scalars["PTSPHY"], dtype=np.float32)
    for name, bar in (("PFPLSN", 2e-2), ("PFHPSN", 2e-2), ("PFSQIF", 1e-1), ("PFSQLF", 3e-1), ("PCOVPTOT", 1e-3),
                      ("PLUDE", 1e-4), ("TENDENCY_LOC_A", 1e-2)):
        a, b = A[name], B[name].astype(np.float64)
        assert np.isfinite(b).all()
        assert np.abs(a - b).sum() <= bar * np.abs(a).sum(), name


def test_oracle_is_not_interposed_by_the_product_library(built, input_data):
    """libcloudsc_b200.so exports the drop-in symbols cloudsc_c / yrecldp / rg ...; the checkers are
    linked -Bsymbolic so that they keep using their OWN kernel and parameters even when the product
    library is loaded first (otherwise parity would compare the GPU path with itself)."""
    import ctypes
    import pyoracle
    from dwarf_p_cloudsc_b200 import api, params, state   # product library loaded BEFORE the oracle
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    ref = state.load_reference()
    for orc in _oracles(built):
        own = ctypes.cast(getattr(orc.lib, "cloudsc_c" if orc.kind == "reference" else "cloudsc_oracle"), ctypes.c_void_p).value
        prod = ctypes.cast(api.lib.cloudsc_c, ctypes.c_void_p).value
        assert own != prod
        F = state.blocked_fields(cols, 100, 32, klev)
        orc.run(F, p, k, 100, 32, klev, scalars["PTSPHY"], 2)
        got = state.unblock(F["PFPLSN"], 100)
        assert np.abs(got - ref["PFPLSN"]).max() <= 1e-12 * np.abs(ref["PFPLSN"]).max()


AEROSOL_CASES = [(1, 0), (0, 1), (1, 1)]


@pytest.mark.parametrize("sed,auto", AEROSOL_CASES)
@pytest.mark.parametrize("variant", ["aer", "aer+warm+land+conv+rain+moist+tend", "aer+sparse"])
def test_aerosol_switches_are_bit_identical_on_the_cpu(built, input_data, sed, auto, variant):
    """yrecldp->laericesed (cloudsc_c.c:1561-1566, reads pre_ice) and laericeauto (:1625-1632, reads picrit_aer,
    pnice) are runtime switches of the reference kernel.  The unmodified reference build, the plain-C port and the
    strict host build of the device column code (its AER instantiation) agree bit for bit; the switches change the
    answer; the fast arithmetic mode stays inside the FP64 tolerance."""
    import pyemul
    import pyoracle
    from dwarf_p_cloudsc_b200 import params, state, validate
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    c2 = state.perturb(cols, variant)
    ngptot, nproma = 100, 16
    A = state.blocked_fields(c2, ngptot, nproma, klev, aerosol=True)
    base = {n: a.copy() for n, a in A.items()}
    orc = pyoracle.best_oracle()
    orc.run(base, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 1)          # switches off
    p.laericesed, p.laericeauto = sed, auto
    runs = {}
    for how in ("best", "port", "emul", "fast"):
        B = {n: a.copy() for n, a in A.items()}
        if how == "best":
            orc.run(B, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 1)
        elif how == "port":
            pyoracle.PortOracle().run(B, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 1)
        else:
            pyemul.Emul(fast=(how == "fast")).run(B, p, k, ngptot, nproma, klev, scalars["PTSPHY"])
        runs[how] = B
    changed = False
    for f in params.OUTPUT_FIELDS:
        ref = state.unblock(runs["best"][f.name], ngptot)
        assert np.isfinite(ref).all(), f.name
        changed |= not np.array_equal(ref, state.unblock(base[f.name], ngptot))
        for how in ("port", "emul"):
            got = state.unblock(runs[how][f.name], ngptot)
            assert np.array_equal(ref, got), (how, f.name, float(np.abs(ref - got).max()))
    assert changed, "the aerosol switches had no effect on this input set"
    errs = validate.validate(runs["fast"], runs["best"], ngptot)
    assert validate.all_pass(errs), "\n" + validate.format_table(errs)


def test_warm_rain_only_aerosol_switches_are_refused(built, input_data):
    """laerliqautolsp / laerliqcoll only act inside the iwarmrain == 1 branch the reference hard-codes away
    (cloudsc_c.c:367,1654,1679); the port and the device column code refuse them instead of silently ignoring them."""
    import pyemul
    from dwarf_p_cloudsc_b200 import params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    F = state.blocked_fields(cols, 32, 32, klev)
    for name in ("laerliqautolsp", "laerliqcoll"):
        setattr(p, name, 1)
        with pytest.raises(RuntimeError):
            pyemul.Emul().run(F, p, k, 32, 32, klev, scalars["PTSPHY"])
        setattr(p, name, 0)


def test_fast_reference_build_is_the_same_kernel(built, input_data):
    """bench.py's CPU arm also times the unmodified reference kernel built for speed (-O3 -march=x86-64-v3/v4, FP
    contraction on; oracle/Makefile `reffast`).  It is a timing build only, but it must still be the same computation:
    within the FP64 tolerance of the -O2 parity build on the shipped and the busiest synthetic input set."""
    import pyoracle
    from dwarf_p_cloudsc_b200 import params, state, validate
    try:
        ref, fast = pyoracle.RefOracle(), pyoracle.RefFastOracle()
    except (FileNotFoundError, OSError) as exc:
        pytest.skip(f"needs oracle/_ref builds this host can run: {exc}")
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    for variant in ("", "warm+land+conv+rain+moist+tend"):
        c2 = state.perturb(cols, variant)
        A = state.blocked_fields(c2, 300, 32, klev)
        B = {n: a.copy() for n, a in A.items()}
        ref.run(A, p, k, 300, 32, klev, scalars["PTSPHY"], 2)
        fast.run(B, p, k, 300, 32, klev, scalars["PTSPHY"], 2)
        errs = validate.validate(B, A, 300)
        assert validate.all_pass(errs), (variant, "\n" + validate.format_table(errs))


@pytest.mark.parametrize("cuts", [(82, 137), (14, 82), (1, 2), (60, 61), (136, 137), (15, 15)])
@pytest.mark.parametrize("variant", ["", "warm+land+conv+rain+moist+tend", "sparse+warm+conv"])
def test_cut_sweep_is_bit_identical(built, oracle, input_data, variant, cuts):
    """The persistent kernel runs the levels of a column tile in two pieces (bandwidth-bound upper part, latency-bound
    cloudy part) with other tiles in between; the state at the cut travels through a buffer.  Host build of exactly
    that code path (IO policy with kRanged, scratch poisoned between the pieces), cut at two arbitrary levels: strict
    mode bit-identical to the reference, fast mode bit-identical to its own uncut sweep."""
    import pyemul
    from dwarf_p_cloudsc_b200 import params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    c2 = state.perturb(cols, variant)
    ngptot, nproma = 100, 16
    A = state.blocked_fields(c2, ngptot, nproma, klev)
    oracle_out = {n: a.copy() for n, a in A.items()}
    oracle.run(oracle_out, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 1)
    for fast in (False, True):
        em = pyemul.Emul(fast=fast)
        whole = {n: a.copy() for n, a in A.items()}
        em.run(whole, p, k, ngptot, nproma, klev, scalars["PTSPHY"], parked=True)
        cut = {n: a.copy() for n, a in A.items()}
        em.run(cut, p, k, ngptot, nproma, klev, scalars["PTSPHY"], cuts=cuts)
        for f in params.OUTPUT_FIELDS:
            a, b = state.unblock(whole[f.name], ngptot), state.unblock(cut[f.name], ngptot)
            assert np.array_equal(a, b), (fast, f.name, float(np.abs(a - b).max()))
            if not fast:
                assert np.array_equal(state.unblock(oracle_out[f.name], ngptot), b), f.name


@pytest.mark.parametrize("ntiles,nsm,per_sm,target_b,ksplit,klev", [(5120, 148, 12, 0, 0, 137), (1792, 148, 12, 0, 0, 137), (1777, 148, 12, 5, 82, 137),
                                                                    (20480, 148, 12, 0, 0, 137), (20480, 148, 12, 12, 136, 137), (7, 2, 3, 0, 0, 137),
                                                                    (5, 2, 3, 3, 1, 3), (1, 1, 1, 0, 0, 60), (100000, 148, 16, 16, 0, 91)])
def test_split_tail_plan_covers_every_level_once(built, ntiles, nsm, per_sm, target_b, ksplit, klev):
    """The block plan of the split-tail launch (csrc/cloudsc_split_plan.h: the text the kernel expands, reached here through the
    host emulation library): every (tile, level) belongs to exactly one block; upper pieces are the FIRST blocks and lower
    pieces the LAST ones (so that in dispatch order every upper piece has started before any lower piece is looked at, and the
    launch ends on short pieces); the number of split tiles respects the flags, the scratch and the launch."""
    import pyemul
    em = pyemul.Emul()
    cut = em.split_cut(ksplit, klev)
    assert 1 <= cut <= klev - 1 and (ksplit <= 0 or cut == min(max(ksplit, 1), klev - 1)) and (ksplit > 0 or cut == max(1, min(3 * klev // 5, klev - 1)))
    nsplit = em.split_count(ntiles, nsm, per_sm, target_b)
    per = min(target_b if target_b > 0 else (per_sm * 5 + 6) // 12, per_sm)
    assert nsplit == min(per * nsm, 4096, ntiles)
    assert em.split_count(ntiles, nsm, per_sm, target_b, 64, 1 << 30) <= 64 and em.split_count(ntiles, nsm, per_sm, target_b, 4096, 3) <= 3
    nblocks = ntiles + nsplit
    cover = np.zeros((ntiles, klev), dtype=np.int32)
    upper_at, lower_at = {}, {}
    for b in range(nblocks):
        pc = em.split_piece(b, ntiles, nsplit, cut, klev)
        assert 0 <= pc["tile"] < ntiles and 0 <= pc["kb"] < pc["ke"] <= klev, (b, pc)
        cover[pc["tile"], pc["kb"]:pc["ke"]] += 1
        if pc["upper"]:
            assert not pc["lower"] and (pc["kb"], pc["ke"]) == (0, cut) and b == pc["tile"] < nsplit
            upper_at[pc["tile"]] = b
        elif pc["lower"]:
            assert (pc["kb"], pc["ke"]) == (cut, klev) and b >= ntiles and pc["tile"] == b - ntiles
            lower_at[pc["tile"]] = b
        else:
            assert (pc["kb"], pc["ke"]) == (0, klev) and nsplit <= b < ntiles
    assert (cover == 1).all()
    assert sorted(upper_at) == sorted(lower_at) == list(range(nsplit))
    if nsplit:
        assert max(upper_at.values()) < nsplit <= ntiles <= min(lower_at.values())


@pytest.mark.parametrize("machine", [(2, 3, 0, 0), (1, 2, 2, 136), (3, 2, 1, 1), (2, 12, 0, 40)])
@pytest.mark.parametrize("variant,ngptot,nproma", [("", 300, 32), ("warm+land+conv+rain+moist+tend", 333, 64), ("sparse+warm+conv", 100, 100)])
def test_split_tail_launch_is_bit_identical(built, oracle, input_data, variant, ngptot, nproma, machine):
    """Schedule mode 3 end to end on the host: the blocks of a split-tail launch walked in dispatch order on a pretend machine
    (nsm x per_sm resident CTAs, so that a few hundred columns make several rounds of tiles), every block sweeping its
    32-column tile over its level range with the device column code, parked state through per-tile buffers guarded by the
    flags.  Strict mode: bit-identical to the reference; fast mode: bit-identical to its own uncut sweep.  (Ragged last
    tiles, tiles straddling NPROMA blocks, every tile split, one tile split.)"""
    import pyemul
    from dwarf_p_cloudsc_b200 import params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    c2 = state.perturb(cols, variant) if variant else cols
    A = state.blocked_fields(c2, ngptot, nproma, klev)
    nsm, per_sm, target_b, ksplit = machine
    ntiles = (ngptot + 31) // 32
    oracle_out = {n: a.copy() for n, a in A.items()}
    oracle.run(oracle_out, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 1)
    for fast in (False, True):
        em = pyemul.Emul(fast=fast)
        whole = {n: a.copy() for n, a in A.items()}
        em.run(whole, p, k, ngptot, nproma, klev, scalars["PTSPHY"], parked=True)
        split = {n: a.copy() for n, a in A.items()}
        info = em.run_split_tail(split, p, k, ngptot, nproma, klev, scalars["PTSPHY"], nsm, per_sm, ksplit, target_b)
        assert info["nsplit"] == em.split_count(ntiles, nsm, per_sm, target_b) > 0 and info["blocks"] == ntiles + info["nsplit"]
        assert info["carry_loads"] == info["nsplit"] and info["cut"] == em.split_cut(ksplit, klev)
        for f in params.OUTPUT_FIELDS:
            a, b = state.unblock(whole[f.name], ngptot), state.unblock(split[f.name], ngptot)
            assert np.array_equal(a, b), (fast, f.name, float(np.abs(a - b).max()))
            if not fast:
                assert np.array_equal(state.unblock(oracle_out[f.name], ngptot), b), f.name


def test_fp32_reference_pins_the_single_precision_column_code(built, input_data):
    """oracle/_ref32 = the UNMODIFIED reference kernel re-typed to single precision by the preprocessor (what JPRB=real32
    does to the Fortran kernel).  (1) The strict host build of the device column code instantiated for float is that
    computation: every field within 1e-7 L1-relative of it (not bit-identical: per-run constants such as
    pow(rcl_kk_cloud_num, rcl_kkbaun) are formed in double before narrowing).  (2) The fast arithmetic mode stays inside
    the GPU-vs-CPU FP32 bars.  (3) The reference's own FP32 build is as far from its FP64 self as validate.FP32_L1_BARS
    allows any FP32 result to be -- the size of those bars is a property of the scheme in single precision."""
    import pyemul
    import pyoracle
    from dwarf_p_cloudsc_b200 import params, state, validate
    try:
        r32, r64 = pyoracle.Ref32Oracle(), pyoracle.RefOracle()
    except (FileNotFoundError, OSError):
        pytest.skip("needs oracle/_ref (reference checkout not present and no prebuilt copy)")
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    ngptot, nproma = 300, 100
    worst_vs64 = 0.0
    for variant in ("", "warm+land+conv+rain+moist+tend"):
        src = state.perturb(cols, variant)
        F = state.blocked_fields(src, ngptot, nproma, klev)
        R64 = {n: a.copy() for n, a in F.items()}
        r64.run(R64, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 2)
        F32 = {n: (a.astype(np.float32) if a.dtype == np.float64 else a.copy()) for n, a in F.items()}
        R32 = {n: a.copy() for n, a in F32.items()}
        r32.run(R32, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 2)
        S = {n: a.copy() for n, a in F32.items()}
        pyemul.Emul().run(S, p, k, ngptot, nproma, klev, scalars["PTSPHY"], dtype=np.float32)
        Ff = {n: a.copy() for n, a in F32.items()}
        pyemul.Emul(fast=True).run(Ff, p, k, ngptot, nproma, klev, scalars["PTSPHY"], dtype=np.float32)
        as64 = lambda A: {n: a.astype(np.float64) if a.dtype == np.float32 else a for n, a in A.items()}
        strict = validate.validate(as64(S), as64(R32), ngptot)
        assert max(e.rel_l1 for e in strict) <= 1e-7, (variant, "\n" + validate.format_table(strict, "fp32ref"))
        fast = validate.validate(as64(Ff), as64(R32), ngptot)
        assert validate.all_pass(fast, "fp32ref"), (variant, "\n" + validate.format_table(fast, "fp32ref"))
        vs64 = validate.validate(as64(R32), R64, ngptot)
        assert validate.all_pass(vs64, "fp32"), (variant, "\n" + validate.format_table(vs64, "fp32"))
        worst_vs64 = max(worst_vs64, max(e.rel_l1 for e in vs64))
    assert worst_vs64 > 1e-2      # the inherent single-precision error of the scheme is of the size of the FP32-vs-FP64 bars
