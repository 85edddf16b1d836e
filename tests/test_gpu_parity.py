"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle.

Tolerances (FP64): per field  sum|err|/sum|ref| <= 1e-10  and  max|err| <= 1e-9 * max|ref|
(SURVEY 8c).  Bit-exactness is not attainable (FMA contraction, CUDA libm).
"""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _torch_fields(F, device="cuda"):
    import torch
    return {n: torch.from_numpy(a).to(device) for n, a in F.items()}


def _run_gpu(F, p, k, ngptot, nproma, klev, ptsphy):
    import torch
    from dwarf_p_cloudsc_b200 import api
    api.set_params(p, k)
    D = _torch_fields(F)
    api.launch(D, ngptot, nproma, klev, ptsphy)
    torch.cuda.synchronize()
    return {n: t.cpu().numpy() for n, t in D.items()}


def _check_against_fp32_reference(G, F, p, k, ngptot, nproma, klev, ptsphy):
    """The FP32 device result against the reference's own kernel re-typed to single precision (oracle/_ref32), per-field
    bars of validate.FP32_L1_BARS_VS_REF32; the CPU-FP32 reference runs with NPROMA <= 128 (stack temporaries)."""
    import pyoracle
    from dwarf_p_cloudsc_b200 import state, validate
    try:
        r32 = pyoracle.Ref32Oracle()
    except (FileNotFoundError, OSError):
        return
    onp = min(nproma, 128)
    cols32 = {n: state.unblock(a, ngptot) for n, a in F.items()}
    R32 = {n: state.expand(a.astype(np.float32) if a.dtype == np.float64 else a, ngptot, onp) for n, a in cols32.items()}
    r32.run(R32, p, k, ngptot, onp, klev, ptsphy, 0)
    errs = [validate.field_error(n, state.unblock(np.asarray(G[n], dtype=np.float64), ngptot),
                                 state.unblock(R32[n].astype(np.float64), ngptot), ngptot) for n in state.VALIDATE_ORDER]
    assert validate.all_pass(errs, "fp32ref"), "\n" + validate.format_table(errs, "fp32ref")


def _compare(G, R, ngptot, precision="fp64"):
    from dwarf_p_cloudsc_b200 import validate
    errs = validate.validate(G, R, ngptot)
    table = validate.format_table(errs, precision)
    assert validate.all_pass(errs, precision), "\n" + table
    return errs


@pytest.mark.parametrize("ngptot,nproma", [(100, 100), (100, 16), (1000, 128), (257, 32), (16384, 32)])
def test_gpu_matches_oracle_shipped_data(built, oracle, input_data, ngptot, nproma):
    from dwarf_p_cloudsc_b200 import params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    F = state.blocked_fields(cols, ngptot, nproma, klev)
    R = {n: a.copy() for n, a in F.items()}
    oracle.run(R, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 0)
    G = _run_gpu(F, p, k, ngptot, nproma, klev, scalars["PTSPHY"])
    _compare(G, R, ngptot)


def test_gpu_matches_reference_h5(built, input_data):
    """The dwarf's own check: 100 columns vs reference.h5 (cloudsc_validate)."""
    from dwarf_p_cloudsc_b200 import params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    F = state.blocked_fields(cols, 100, 100, klev)
    G = _run_gpu(F, p, k, 100, 100, klev, scalars["PTSPHY"])
    ref = state.load_reference()
    R = {n: state.expand(a, 100, 100) for n, a in ref.items()}
    _compare(G, R, 100)


@pytest.mark.parametrize("variant,nssopt", [("warm", 1), ("hot+land", 1), ("conv+land", 1), ("rain+warm", 1),
                                            ("moist+tend", 0), ("moist+tend", 2), ("moist+tend", 3),
                                            ("warm+land+conv+rain+moist+tend", 1), ("hot+rain+moist+conv", 1),
                                            ("sparse", 1), ("sparse", 0), ("sparse+warm+conv", 1)])
def test_gpu_matches_oracle_perturbed(built, oracle, input_data, variant, nssopt):
    """Branches the shipped data never takes (warm phase, land, snow detrainment, rain freezing...)."""
    from dwarf_p_cloudsc_b200 import params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    p.nssopt = nssopt
    c2 = state.perturb(cols, variant)
    ngptot, nproma = 400, 64
    F = state.blocked_fields(c2, ngptot, nproma, klev)
    R = {n: a.copy() for n, a in F.items()}
    oracle.run(R, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 0)
    G = _run_gpu(F, p, k, ngptot, nproma, klev, scalars["PTSPHY"])
    _compare(G, R, ngptot)


def test_full_size_properties(built, input_data):
    """C2 size (163840 x 128): replication invariance -- column g must equal column g mod 100 --
    and agreement of the first 100 columns with reference.h5."""
    import torch
    from dwarf_p_cloudsc_b200 import params, state, api, validate
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    ngptot, nproma = 163840, 128
    F = state.blocked_fields(cols, ngptot, nproma, klev)
    G = _run_gpu(F, p, k, ngptot, nproma, klev, scalars["PTSPHY"])
    ref = state.load_reference()
    for name in state.VALIDATE_ORDER:
        u = state.unblock(G[name], ngptot)
        base = u[..., :100]
        idx = np.arange(ngptot) % 100
        assert np.array_equal(u, base[..., idx]), f"{name}: replicated columns differ"
        e = validate.field_error(name, base, ref[name], 100)
        assert e.passes(), (name, e)


def _host_arrays(F, pinned):
    import torch
    H = {n: torch.from_numpy(a.copy()) for n, a in F.items()}
    return {n: (t.pin_memory() if pinned else t) for n, t in H.items()}


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("variant,ngptot,nproma,chunk,lanes", [("", 5000, 64, 7, 3), ("warm+land+conv+rain+moist+tend", 4133, 128, 5, 4),
                                                               ("sparse+warm+conv", 1221, 32, 0, 0), ("", 777, 96, 3, 2)])
def test_run_host_matches_oracle(built, oracle, input_data, pinned, variant, ngptot, nproma, chunk, lanes):
    """cloudsc_b200_run_host (the path behind bench.py's `e2e`) against the CPU oracle: pinned and pageable host
    arrays, shipped and perturbed inputs, ragged last block / ragged last chunk, NPROMA not a multiple of 32.
    Also the driver contract around the kernel: species 5 of TENDENCY_LOC_CLD comes back zero and the padding
    columns of a ragged last block keep the caller's values (zeros for PCOVPTOT / TENDENCY_LOC_CLD,
    cloudsc_driver.c:199-200)."""
    from dwarf_p_cloudsc_b200 import api, params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    src = state.perturb(cols, variant) if variant else cols
    F = state.blocked_fields(src, ngptot, nproma, klev)
    R = {n: a.copy() for n, a in F.items()}
    oracle.run(R, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 0)
    for f in params.OUTPUT_FIELDS:
        if f.use == "W":
            F[f.name][...] = -777.0
    api.set_params(p, k)
    H = _host_arrays(F, pinned)
    st = api.run_host(H, ngptot, nproma, klev, scalars["PTSPHY"], chunk_blocks=chunk, nstreams=lanes)
    assert st["h2d_bytes"] > 0 and st["d2h_bytes"] > 0
    G = {n: t.numpy() for n, t in H.items()}
    _compare(G, R, ngptot)
    assert np.all(G["TENDENCY_LOC_CLD"][:, 4] == 0.0)
    pad = ngptot % nproma
    if pad:
        for f in params.OUTPUT_FIELDS:
            tail = G[f.name][-1][..., pad:]
            if f.name in ("PCOVPTOT", "TENDENCY_LOC_CLD"):
                assert np.all(tail == 0.0), f.name
            elif f.use == "W":
                assert np.all(tail == -777.0), f.name
            else:                       # PLUDE is in/out: the padding columns keep the input
                assert np.array_equal(tail, F[f.name][-1][..., pad:]), f.name


def test_run_host_moves_only_live_bytes(built, input_data):
    """What travels is what the kernel reads and writes: 4 of 5 species, no dead upper-level rows
    (include/cloudsc_b200.h).  Counted by the library from the copies it issues."""
    from dwarf_p_cloudsc_b200 import api, params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    ngptot, nproma = 2048, 128
    F = state.blocked_fields(cols, ngptot, nproma, klev)
    api.set_params(p, k)
    st = api.run_host(_host_arrays(F, True), ngptot, nproma, klev, scalars["PTSPHY"])
    ktop = p.ncldtop - 1
    rows_in = 18 * klev + (klev + 1) + 8 * klev - (9 * ktop + (ktop - 1) + (ktop + 1))
    assert st["h2d_bytes"] == ngptot * (rows_in * 8 + 8 + 4)                       # + PLSM + KTYPE
    rows_out = 5 * klev + 4 * klev + 14 * (klev + 1)
    assert st["d2h_bytes"] == ngptot * (rows_out * 8 + 8)                          # + PRAINFRAC_TOPRFZ
    assert st["h2d_bytes"] / ngptot < params.algorithmic_bytes_per_column(klev) - rows_out * 8


def test_skipped_input_rows_are_dead(built, input_data):
    """The rows run_host never uploads (CopyPlan in csrc/cloudsc_abi.cu) are rows no output depends on: poisoning them
    with NaN on the device-resident path leaves every output bit unchanged."""
    from dwarf_p_cloudsc_b200 import params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    ngptot, nproma = 1000, 128
    src = state.perturb(cols, "warm+land+conv+rain+moist+tend")
    F = state.blocked_fields(src, ngptot, nproma, klev)
    G1 = _run_gpu({n: a.copy() for n, a in F.items()}, p, k, ngptot, nproma, klev, scalars["PTSPHY"])
    ktop = p.ncldtop - 1
    P = {n: a.copy() for n, a in F.items()}
    for n in ("PQ", "TENDENCY_TMP_Q", "PSUPSAT", "PSNDE", "PHRSW", "PHRLW", "PVERVEL", "PMFU", "PMFD"):
        P[n][:, :ktop] = np.nan
    P["PAP"][:, :ktop - 1] = np.nan
    P["PLU"][:, :ktop + 1] = np.nan
    P["PCLV"][:, 4] = np.nan
    P["TENDENCY_TMP_CLD"][:, 4] = np.nan
    G2 = _run_gpu(P, p, k, ngptot, nproma, klev, scalars["PTSPHY"])
    for f in params.OUTPUT_FIELDS:
        assert np.array_equal(G1[f.name], G2[f.name]), f.name


def test_run_host_staging_arenas_follow_the_request(built, oracle, input_data):
    """The staging lanes are cached between calls.  A lane first allocated for a small request must be replaced when
    a later call needs more (2 lanes x big chunks, then 4 lanes x small chunks, then 4 lanes x medium chunks --
    lanes 2 and 3 were once left with the small arenas); every call is checked against the oracle, and a guard zone
    cannot exist inside cudaMalloc'ed lanes, so the medium call runs twice and must reproduce itself."""
    from dwarf_p_cloudsc_b200 import api, params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    ngptot, nproma = 6000, 64
    F = state.blocked_fields(cols, ngptot, nproma, klev)
    R = {n: a.copy() for n, a in F.items()}
    oracle.run(R, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 0)
    api.set_params(p, k)
    api.release_cache()
    outs = []
    for chunk, lanes in ((40, 2), (6, 4), (20, 4), (20, 4)):
        H = _host_arrays(F, True)
        api.run_host(H, ngptot, nproma, klev, scalars["PTSPHY"], chunk_blocks=chunk, nstreams=lanes)
        G = {n: t.numpy() for n, t in H.items()}
        _compare(G, R, ngptot)
        outs.append(G)
    for f in params.OUTPUT_FIELDS:
        assert np.array_equal(outs[2][f.name], outs[3][f.name]), f.name
        assert np.array_equal(outs[0][f.name], outs[1][f.name]), f.name


def test_run_host_fp32(built, oracle, input_data):
    """cloudsc_b200_run_host_ex(FP32): float host arrays through the same pipeline == the FP32 resident state."""
    import torch
    from dwarf_p_cloudsc_b200 import api, params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    ngptot, nproma = 3000, 128
    F = state.blocked_fields(cols, ngptot, nproma, klev)
    F32 = {n: (a.astype(np.float32) if a.dtype == np.float64 else a.copy()) for n, a in F.items()}
    api.set_params(p, k)
    H = _host_arrays(F32, True)
    api.run_host(H, ngptot, nproma, klev, scalars["PTSPHY"], chunk_blocks=5, nstreams=3, precision=api.FP32)
    S = {n: a.copy() for n, a in F32.items()}
    st = api.State(ngptot, nproma, klev, api.FP32)
    try:
        st.upload(S); st.step(scalars["PTSPHY"]); st.download(S); st.sync()
    finally:
        st.close()
    for f in params.OUTPUT_FIELDS:
        assert np.array_equal(state.unblock(H[f.name].numpy(), ngptot), state.unblock(S[f.name], ngptot)), f.name


def test_launch_adopts_regularly_spaced_caller_fields(built, input_data):
    """cloudsc_b200_launch() on caller-owned device fields uses the tensor-map kernel when the caller's fields have the
    arena's regularity (here: carved from one allocation in the library's group order, via a state's own pointers) and
    the bulk-copy kernel otherwise; both give the same bits."""
    import torch
    from dwarf_p_cloudsc_b200 import api, params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    ngptot, nproma = 4096, 128
    F = state.blocked_fields(cols, ngptot, nproma, klev)
    api.set_params(p, k)
    c0 = api.launch_counts()
    G1 = _run_gpu({n: a.copy() for n, a in F.items()}, p, k, ngptot, nproma, klev, scalars["PTSPHY"])   # separate allocations
    c1 = api.launch_counts()
    assert c1["tile"] == c0["tile"] + 1 and c1["tmap"] == c0["tmap"]
    st = api.State(ngptot, nproma, klev)
    try:
        st.upload(F)
        ptrs = {f.name: st.field_ptr(f.name) for f in params.FIELDS if st.field_ptr(f.name)}
        api.launch(ptrs, ngptot, nproma, klev, scalars["PTSPHY"])       # the state's fields, handed over as raw pointers
        c2 = api.launch_counts()
        assert c2["tmap"] == c1["tmap"] + 1 and c2["tile"] == c1["tile"]
        G2 = {n: a.copy() for n, a in F.items()}
        st.download(G2); st.sync()
    finally:
        st.close()
    for f in params.OUTPUT_FIELDS:
        assert np.array_equal(G1[f.name], G2[f.name]), f.name


@pytest.mark.parametrize("sed,auto", [(1, 0), (0, 1), (1, 1)])
def test_aerosol_switches_match_oracle(built, oracle, input_data, sed, auto):
    """yrecldp->laericesed / laericeauto (cloudsc_c.c:1561-1566, 1625-1632): runtime switches of the reference kernel
    that read PRE_ICE / PICRIT_AER / PNICE.  All four entry points (launch, resident state, run_host, the cloudsc_c()
    shim) against the unmodified reference build; switched off again, the default kernels come back."""
    import torch
    from dwarf_p_cloudsc_b200 import api, params, state, validate
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    p.laericesed, p.laericeauto = sed, auto
    src = state.perturb(cols, "aer+warm+land+conv+rain+moist+tend")
    ngptot, nproma = 777, 128
    F = state.blocked_fields(src, ngptot, nproma, klev, aerosol=True)
    R = {n: a.copy() for n, a in F.items()}
    oracle.run(R, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 0)
    base = {n: a.copy() for n, a in F.items()}
    p0, _ = params.params_from_scalars(scalars)
    oracle.run(base, p0, k, ngptot, nproma, klev, scalars["PTSPHY"], 0)
    assert not np.array_equal(base["PFPLSN"], R["PFPLSN"]), "the switches must change the answer"
    try:
        c0 = api.launch_counts()
        G = _run_gpu({n: a.copy() for n, a in F.items()}, p, k, ngptot, nproma, klev, scalars["PTSPHY"])
        _compare(G, R, ngptot)
        assert api.launch_counts()["direct_aer"] == c0["direct_aer"] + 1
        # resident state: holds the aerosol inputs because the switches were on when it was created
        S = {n: a.copy() for n, a in F.items()}
        st = api.State(ngptot, nproma, klev)
        try:
            st.upload(S); st.step(scalars["PTSPHY"]); st.download(S); st.sync()
        finally:
            st.close()
        _compare(S, R, ngptot)
        # host pipeline
        H = _host_arrays(F, True)
        api.run_host(H, ngptot, nproma, klev, scalars["PTSPHY"], chunk_blocks=2, nstreams=2)
        _compare({n: t.numpy() for n, t in H.items()}, R, ngptot)
        # drop-in shim, first block
        api.set_globals(p, k)
        B = {n: a.copy() for n, a in F.items()}
        rc = api.cloudsc_c(1, nproma, nproma, klev, scalars["PTSPHY"], {n: a[0] for n, a in B.items()})
        assert rc == 0, api.lib.cloudsc_b200_last_error()
        for f in params.OUTPUT_FIELDS:
            e = validate.field_error(f.name, B[f.name][0], R[f.name][0], nproma)
            assert e.passes(), (f.name, e)
        # a state created without the switches cannot serve them
        api.set_params(p0, k)
        st = api.State(ngptot, nproma, klev)
        try:
            api.set_params(p, k)
            with pytest.raises(api.CloudscError):
                st.step(scalars["PTSPHY"])
        finally:
            st.close()
    finally:
        api.set_params(p0, k)
        api.set_globals(p0, k)
        api.release_thread()
    G0 = _run_gpu({n: a.copy() for n, a in F.items() if n not in state.AEROSOL_FIELDS}, p0, k, ngptot, nproma, klev, scalars["PTSPHY"])
    _compare(G0, base, ngptot)


def test_run_host_matches_resident(built, input_data):
    """End-to-end host pipeline (chunked, multi-stream) == single resident launch, bit for bit."""
    import torch
    from dwarf_p_cloudsc_b200 import params, state, api
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    ngptot, nproma = 5000, 64
    F = state.blocked_fields(cols, ngptot, nproma, klev)
    G = _run_gpu({n: a.copy() for n, a in F.items()}, p, k, ngptot, nproma, klev, scalars["PTSPHY"])
    H = {n: torch.from_numpy(a.copy()).pin_memory() for n, a in F.items()}
    st = api.run_host(H, ngptot, nproma, klev, scalars["PTSPHY"], chunk_blocks=7, nstreams=3)
    assert st["h2d_bytes"] > 0 and st["d2h_bytes"] > 0
    for f in params.OUTPUT_FIELDS:
        assert np.array_equal(state.unblock(H[f.name].numpy(), ngptot), state.unblock(G[f.name], ngptot)), f.name


def test_cloudsc_c_dropin(built, oracle, input_data):
    """The binary-compatible cloudsc_c(): one block, host pointers, parameters from C globals,
    partial column range -- columns outside kidia..kfdia must stay untouched."""
    from dwarf_p_cloudsc_b200 import params, state, api
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    nproma = 100
    F = state.blocked_fields(cols, 100, nproma, klev)
    R = {n: a.copy() for n, a in F.items()}
    oracle.run(R, p, k, 100, nproma, klev, scalars["PTSPHY"], 1)
    api.set_globals(p, k)
    for f in params.OUTPUT_FIELDS:
        if f.use == "W":
            F[f.name][...] = -777.0
    F["PCOVPTOT"][...] = 0.0
    F["TENDENCY_LOC_CLD"][...] = 0.0
    kidia, kfdia = 11, 90
    rc = api.cloudsc_c(kidia, kfdia, nproma, klev, scalars["PTSPHY"], {n: a[0] for n, a in F.items()})
    assert rc == 0, api.lib.cloudsc_b200_last_error()
    from dwarf_p_cloudsc_b200 import validate
    for f in params.OUTPUT_FIELDS:
        got, ref = F[f.name][0], R[f.name][0]
        inner = validate.field_error(f.name, got[..., kidia - 1:kfdia], ref[..., kidia - 1:kfdia], kfdia - kidia + 1)
        assert inner.passes(), (f.name, inner)
        if f.use == "W" and f.name not in ("PCOVPTOT", "TENDENCY_LOC_CLD"):
            assert np.all(got[..., :kidia - 1] == -777.0) and np.all(got[..., kfdia:] == -777.0), f.name


def test_state_handle_api(built, oracle, input_data):
    import torch
    from dwarf_p_cloudsc_b200 import params, state, api, validate
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    ngptot, nproma = 1000, 128
    F = state.blocked_fields(cols, ngptot, nproma, klev)
    R = {n: a.copy() for n, a in F.items()}
    oracle.run(R, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 0)
    api.set_params(p, k)
    st = api.State(ngptot, nproma, klev)
    st.upload(F)
    st.step(scalars["PTSPHY"])
    st.reset_inout()           # the kernel scales PLUDE in place: a second step needs the pristine copy
    st.step(scalars["PTSPHY"], 0, 3)
    st.step(scalars["PTSPHY"], 3, 5)
    st.download(F)
    st.sync()
    st.close()
    errs = validate.validate(F, R, ngptot)
    assert validate.all_pass(errs), "\n" + validate.format_table(errs)


def test_unsupported_params_fail_loudly(built, input_data):
    from dwarf_p_cloudsc_b200 import params, api
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    p.laerliqautolsp = 1           # only acts in the iwarmrain == 1 branch the reference compiles dead (cloudsc_c.c:367,1654)
    with pytest.raises(api.CloudscError) as ei:
        api.set_params(p, k)
    assert ei.value.code == -1 and "laerliqautolsp" in str(ei.value)
    p.laerliqautolsp, p.ncldtop = 0, 1
    with pytest.raises(api.CloudscError) as ei:
        api.set_params(p, k)
    assert ei.value.code == -1 and "ncldtop" in str(ei.value)
    with pytest.raises(api.CloudscError):
        api.set_variant(3)             # former alias ids are gone


@pytest.mark.parametrize("kernel_variant", [2, 5, 7, 8, 9])
def test_every_kernel_variant_matches_oracle(built, oracle, input_data, kernel_variant):
    """All register-budget / IO-policy variants run the same arithmetic: each must meet the parity bar on
    a ragged geometry (last block partly filled, column count not a multiple of the warp size)."""
    from dwarf_p_cloudsc_b200 import params, state, api
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    c2 = state.perturb(cols, "warm+land+conv+rain+moist+tend")
    try:
        api.set_variant(kernel_variant)
        for src, (ngptot, nproma) in ((cols, (1000, 128)), (cols, (1221, 64)), (c2, (333, 32))):
            F = state.blocked_fields(src, ngptot, nproma, klev)
            R = {n: a.copy() for n, a in F.items()}
            oracle.run(R, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 0)
            G = _run_gpu(F, p, k, ngptot, nproma, klev, scalars["PTSPHY"])
            _compare(G, R, ngptot)
    finally:
        api.set_variant(0)


@pytest.mark.parametrize("args", [("1", "100", "16"), ("1", "16384", "128"), ("1", "1000", "100")])
def test_c_driver_executable_validates_against_reference_h5(built, args):
    """dwarf-cloudsc-b200 NUMOMP NGPTOT NPROMA: the C host (load_state from input.h5, GPU sweep,
    cloudsc_validate against reference.h5) -- the dwarf-cloudsc-c command line end to end."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "dwarf_p_cloudsc_b200", "lib", "dwarf-cloudsc-b200")
    r = subprocess.run([exe, *args], cwd=os.path.join(root, "tests", "golden"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    rows = [l for l in r.stdout.splitlines() if " 2D" in l]
    assert len(rows) == 21 and all(l.rstrip().endswith("PASS") for l in rows), r.stdout
    assert "TOTAL" in r.stdout and f"NGPTOT={args[1]}, NPROMA={args[2]}" in r.stdout


@pytest.mark.parametrize("kernel_variant", [0, 2])
def test_fp32_build_within_the_stated_single_precision_bars(built, oracle, input_data, kernel_variant):
    """BASELINE configs[4]: the FP32 build (float device fields, float arithmetic) against the FP64 oracle,
    per-field bars of validate.FP32_L1_BARS; every value finite."""
    import torch
    from dwarf_p_cloudsc_b200 import api, params, state, validate
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    ngptot, nproma = 1000, 128
    F = state.blocked_fields(cols, ngptot, nproma, klev)
    R = {n: a.copy() for n, a in F.items()}
    oracle.run(R, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 0)
    api.set_params(p, k)
    D = {n: torch.from_numpy(a).cuda() for n, a in F.items()}
    D = {n: (t.float() if t.dtype == torch.float64 else t) for n, t in D.items()}
    try:
        api.set_variant(kernel_variant)
        api.launch(D, ngptot, nproma, klev, scalars["PTSPHY"], precision=api.FP32)
        torch.cuda.synchronize()
    finally:
        api.set_variant(0)
    G = {n: t.double().cpu().numpy() for n, t in D.items()}
    for n in state.VALIDATE_ORDER:
        assert np.isfinite(G[n]).all(), n
    errs = validate.validate(G, R, ngptot)
    assert validate.all_pass(errs, "fp32"), "\n" + validate.format_table(errs, "fp32")
    _check_against_fp32_reference(G, F, p, k, ngptot, nproma, klev, scalars["PTSPHY"])


@pytest.mark.parametrize("ngptot,nproma", [(1000, 128), (4099, 64), (777, 96)])
def test_fp32_resident_state_within_the_single_precision_bars(built, oracle, input_data, ngptot, nproma):
    """FP32 resident state (cloudsc_b200_create(..., FP32): float host arrays, float arena, FLOAT32 tensor maps,
    16 resident warps per SM) against the FP64 oracle within validate.FP32_L1_BARS, incl. a ragged tail; and it
    agrees with cloudsc_b200_launch(FP32) on the same float fields to rounding-level differences."""
    import torch
    from dwarf_p_cloudsc_b200 import api, params, state, validate
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    F = state.blocked_fields(cols, ngptot, nproma, klev)
    R = {n: a.copy() for n, a in F.items()}
    oracle.run(R, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 0)
    api.set_params(p, k)
    H = {n: (a.astype(np.float32) if a.dtype == np.float64 else a.copy()) for n, a in F.items()}
    st = api.State(ngptot, nproma, klev, api.FP32)
    try:
        info = st.kernel_info()
        assert info["variant"] == 100 and info["local_bytes"] == 0 and info["blocks_per_sm"] >= 16, info
        st.upload(H)
        st.step(scalars["PTSPHY"])
        st.download(H)
        st.sync()
    finally:
        st.close()
    G = {n: a.astype(np.float64) if a.dtype == np.float32 else a for n, a in H.items()}
    for n in state.VALIDATE_ORDER:
        assert np.isfinite(G[n]).all(), n
    errs = validate.validate(G, R, ngptot)
    assert validate.all_pass(errs, "fp32"), "\n" + validate.format_table(errs, "fp32")
    _check_against_fp32_reference(G, F, p, k, ngptot, nproma, klev, scalars["PTSPHY"])
    # same arithmetic through caller-owned float fields
    D = {n: torch.from_numpy(a.astype(np.float32) if a.dtype == np.float64 else a).cuda() for n, a in F.items()}
    api.launch(D, ngptot, nproma, klev, scalars["PTSPHY"], precision=api.FP32)
    torch.cuda.synchronize()
    for n in state.VALIDATE_ORDER:
        a = state.unblock(D[n].cpu().numpy(), ngptot)[..., :ngptot]
        b = state.unblock(H[n], ngptot)[..., :ngptot]
        assert np.array_equal(a, b), n


@pytest.mark.parametrize("kernel_variant", [0, 2, 5, 8, 9])
def test_no_write_outside_the_fields(built, input_data, kernel_variant):
    """Guard zones around every device field (compute-sanitizer is not available on the pool): 4 KB of a
    canary pattern before and after each field must survive a launch on ragged geometries, and input
    fields must come back unchanged (PLUDE is the only in/out field)."""
    import torch
    from dwarf_p_cloudsc_b200 import api, params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    api.set_params(p, k)
    guard = 512                                    # elements of 8 bytes
    canary = -1.2345678e300
    try:
        api.set_variant(kernel_variant)
        for ngptot, nproma in ((257, 32), (300, 128), (77, 16), (1221, 64)):
            F = state.blocked_fields(cols, ngptot, nproma, klev)
            D, arenas = {}, {}
            for n, a in F.items():
                flat = torch.from_numpy(a.reshape(-1).view(np.float64 if a.dtype == np.float64 else np.int32))
                if a.dtype == np.int32:             # KTYPE: int field, pad in int32 units
                    ar = torch.full((flat.numel() + 4 * guard,), 0x7f7f7f7f, dtype=torch.int32, device="cuda")
                    ar[2 * guard:2 * guard + flat.numel()] = flat.cuda()
                    D[n] = ar[2 * guard:2 * guard + flat.numel()]
                else:
                    ar = torch.full((flat.numel() + 2 * guard,), canary, dtype=torch.float64, device="cuda")
                    ar[guard:guard + flat.numel()] = flat.cuda()
                    D[n] = ar[guard:guard + flat.numel()]
                arenas[n] = ar
            api.launch(D, ngptot, nproma, klev, scalars["PTSPHY"])
            torch.cuda.synchronize()
            for n, ar in arenas.items():
                if ar.dtype == torch.int32:
                    assert bool((ar[:2 * guard] == 0x7f7f7f7f).all()) and bool((ar[-2 * guard:] == 0x7f7f7f7f).all()), n
                else:
                    assert bool((ar[:guard] == canary).all()) and bool((ar[-guard:] == canary).all()), (n, ngptot, nproma)
            for f in params.INPUT_FIELDS:
                if f.use == "R" and f.name in F:
                    assert np.array_equal(D[f.name].cpu().numpy().view(F[f.name].dtype).reshape(F[f.name].shape), F[f.name]), f.name
    finally:
        api.set_variant(0)


def test_columns_are_independent(built, input_data):
    """Permuting the source columns permutes the outputs bit for bit: no value leaks between the columns
    of a warp / CTA through the shared-memory staging."""
    import torch
    from dwarf_p_cloudsc_b200 import params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    rng = np.random.default_rng(7)
    perm = rng.permutation(100)
    c2 = {n: np.ascontiguousarray(a[..., perm]) for n, a in cols.items()}
    ngptot, nproma = 1000, 128
    G1 = _run_gpu(state.blocked_fields(cols, ngptot, nproma, klev), p, k, ngptot, nproma, klev, scalars["PTSPHY"])
    G2 = _run_gpu(state.blocked_fields(c2, ngptot, nproma, klev), p, k, ngptot, nproma, klev, scalars["PTSPHY"])
    idx = perm[np.arange(ngptot) % 100]            # column g of run 2 holds source column perm[g mod 100]
    for f in params.OUTPUT_FIELDS:
        u1 = state.unblock(G1[f.name], ngptot)[..., :100]
        u2 = state.unblock(G2[f.name], ngptot)
        assert np.array_equal(u2, u1[..., idx]), f.name


@pytest.mark.parametrize("ngptot,nproma,perturbation", [(1000, 128, ""), (257, 32, ""), (1221, 64, "warm+land+conv+rain+moist+tend"),
                                                        (4096, 256, "hot+rain+moist+conv"), (100, 100, "")])
def test_arena_state_tensor_map_kernel_matches_oracle(built, oracle, input_data, ngptot, nproma, perturbation):
    """cloudsc_b200_create/upload/step/download: the arena-backed state whose step() fetches a level's
    inputs with 4 tensor-map TMA requests (NPROMA multiple of 32; plain kernel otherwise), in block ranges."""
    from dwarf_p_cloudsc_b200 import params, state, api, validate
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    src = state.perturb(cols, perturbation) if perturbation else cols
    F = state.blocked_fields(src, ngptot, nproma, klev)
    # the CPU reference keeps ~51 KB of stack temporaries per column of a block: run it with NPROMA <= 128
    # and compare column by column (the result of a column does not depend on the blocking)
    onp = min(nproma, 128)
    R = state.blocked_fields(src, ngptot, onp, klev)
    oracle.run(R, p, k, ngptot, onp, klev, scalars["PTSPHY"], 0)
    api.set_params(p, k)
    st = api.State(ngptot, nproma, klev)
    try:
        info = st.kernel_info()
        assert (info["variant"] == 100) == (nproma % 32 == 0), info      # 100 = the tensor-map kernel, one CTA per tile
        assert st.get_schedule()["used"] == 0                            # (a launch this small is never split)
        st.upload(F)
        nb = state.nblocks_of(ngptot, nproma)
        half = nb // 2
        if half:
            st.step(scalars["PTSPHY"], 0, half)
        st.step(scalars["PTSPHY"], half, nb - half)
        st.download(F)
        st.sync()
    finally:
        st.close()
    errs = [validate.field_error(n, state.unblock(F[n], ngptot), state.unblock(R[n], ngptot), ngptot) for n in state.VALIDATE_ORDER]
    assert validate.all_pass(errs), "\n" + validate.format_table(errs)


@pytest.mark.parametrize("lev0,lev1", [(97, 137), (0, 30), (60, 63)])
def test_other_column_heights(built, oracle, input_data, lev0, lev1):
    """Nothing is specialised on KLEV = 137: a slab of the levels (with its own half-level pressures) is a
    valid, shorter column set for the reference kernel and for both device paths."""
    import torch
    from dwarf_p_cloudsc_b200 import params, state, api, validate
    cols, scalars, klon, _ = input_data
    p, k = params.params_from_scalars(scalars)
    if lev1 - lev0 < 16:
        p.ncldtop = 2
    klev = lev1 - lev0
    sub = {}
    for f in params.FIELDS:
        if f.name not in cols:
            continue
        a = cols[f.name]
        if f.kind == "L":
            sub[f.name] = np.ascontiguousarray(a[lev0:lev1])
        elif f.kind == "H":
            sub[f.name] = np.ascontiguousarray(a[lev0:lev1 + 1])
        elif f.kind == "C":
            sub[f.name] = np.ascontiguousarray(a[:, lev0:lev1])
        else:
            sub[f.name] = a
    ngptot, nproma = 700, 64
    F = state.blocked_fields(sub, ngptot, nproma, klev)
    R = {n: a.copy() for n, a in F.items()}
    oracle.run(R, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 0)
    api.set_params(p, k)
    G = {n: a.copy() for n, a in F.items()}
    st = api.State(ngptot, nproma, klev)
    try:
        st.upload(G); st.step(scalars["PTSPHY"]); st.download(G); st.sync()
    finally:
        st.close()
    errs = validate.validate(G, R, ngptot)
    assert validate.all_pass(errs), "\n" + validate.format_table(errs)
    D = {n: torch.from_numpy(a).cuda() for n, a in F.items()}
    api.launch(D, ngptot, nproma, klev, scalars["PTSPHY"])
    torch.cuda.synchronize()
    errs = validate.validate({n: t.cpu().numpy() for n, t in D.items()}, R, ngptot)
    assert validate.all_pass(errs), "\n" + validate.format_table(errs)


def test_full_size_properties_arena_state(built, input_data):
    """The path bench.py times, at its size (163840 x 128, resident arena state, tensor-map kernel in its default schedule):
    replication invariance (column g == column g mod 100, bitwise), agreement of the first 100 columns with
    reference.h5, and bitwise agreement with the kernel behind cloudsc_b200_launch on caller-owned fields."""
    from dwarf_p_cloudsc_b200 import params, state, api, validate
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    ngptot, nproma = 163840, 128
    F = state.blocked_fields(cols, ngptot, nproma, klev)
    G2 = _run_gpu({n: a.copy() for n, a in F.items()}, p, k, ngptot, nproma, klev, scalars["PTSPHY"])
    st = api.State(ngptot, nproma, klev)
    try:
        # an FP64 state of this size gets the split-tail launch (schedule mode 3) without being asked
        sched, info = st.get_schedule(), st.kernel_info()
        assert sched["mode"] == 3 and sched["used"] == 3 and info["variant"] == 103 and info["local_bytes"] == 0, (sched, info)
        st.upload(F); st.step(scalars["PTSPHY"]); st.download(F); st.sync()
    finally:
        st.close()
    ref = state.load_reference()
    idx = np.arange(ngptot) % 100
    for name in state.VALIDATE_ORDER:
        u = state.unblock(F[name], ngptot)
        base = u[..., :100]
        assert np.array_equal(u, base[..., idx]), f"{name}: replicated columns differ"
        assert validate.field_error(name, base, ref[name], 100).passes(), name
        assert np.array_equal(u, state.unblock(G2[name], ngptot)), f"{name}: the two device paths differ"


@pytest.mark.parametrize("which", ["kcaching", "scc"])
def test_reference_cuda_kernels_agree_with_ours(built, oracle, input_data, which):
    """The reference's own CUDA kernels (cloudsc_c_k_caching.cu / cloudsc_c.cu, unmodified, compiled for sm_100a into
    oracle/_ref by `make -C oracle gpuref`) launched as cloudsc_driver.cu:391-416 does: their outputs meet the FP64 bar
    against the CPU oracle and against the product kernel on the same device fields.  This is the harness behind
    bench.py's `gpu_reference` (the same-device number to beat)."""
    import torch
    import pyoracle
    from dwarf_p_cloudsc_b200 import params, state, validate
    try:
        g = pyoracle.GpuRef(which)
    except FileNotFoundError:
        pytest.skip("oracle/_ref/libcloudsc_gpuref_*.so did not travel with this snapshot")
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    ngptot, nproma = 1024, 128
    F = state.blocked_fields(cols, ngptot, nproma, klev)
    R = {n: a.copy() for n, a in F.items()}
    oracle.run(R, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 0)
    ours = _run_gpu({n: a.copy() for n, a in F.items()}, p, k, ngptot, nproma, klev, scalars["PTSPHY"])
    D = _torch_fields(F)
    t = g.run(D, p, k, ngptot, nproma, scalars["PTSPHY"])
    torch.cuda.synchronize()
    assert t["ms"] > 0
    G = {n: x.cpu().numpy() for n, x in D.items()}
    _compare(G, R, ngptot)
    _compare(G, ours, ngptot)
    info = g.kernel_info()
    assert info["registers"] > 0


@pytest.mark.parametrize("args", [("4", "1000", "128"), ("4", "100", "16"), ("3", "2500", "64")])
def test_level0_link_of_the_reference_driver(built, args):
    """INTEGRATION.md level 0: the reference's OWN dwarf_cloudsc.c + cloudsc_driver.c + cloudsc_validate.c + mycpu.c,
    unmodified, linked against libcloudsc_b200.so (oracle/Makefile `level0`).  Its OpenMP threads call the cloudsc_c()
    shim concurrently on disjoint blocks (cloudsc_driver.c:191-217); the table it prints is the reference's own
    (cloudsc_validate.c:193-216).  Bar per field: AbsMaxErr <= 1e-9 * max|value| (the FP64 max-norm bar; the reference's
    table does not carry sum|ref|), every value finite."""
    import math
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "dwarf-cloudsc-c-level0")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dwarf-cloudsc-c-level0 did not travel with this snapshot")
    env = dict(os.environ, OMP_NUM_THREADS=args[0], OMP_STACKSIZE="64M")
    r = subprocess.run([exe, *args], cwd=os.path.join(root, "tests", "golden"), capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert f"NUMOMP={args[0]}, NGPTOT={args[1]}, NPROMA={args[2]}" in r.stdout, r.stdout
    rows = [l.split() for l in r.stdout.splitlines() if len(l.split()) >= 7 and l.split()[1][:2] == "2D"]
    assert len(rows) == 21, r.stdout
    for row in rows:
        name, vmin, vmax, maxerr, avgerr, relerr = row[0], *map(float, row[2:7])
        assert all(math.isfinite(v) for v in (vmin, vmax, maxerr, avgerr, relerr)), row
        assert maxerr <= 1e-9 * max(abs(vmin), abs(vmax)), (name, maxerr, vmin, vmax, r.stdout)
    # per-thread rows: every OpenMP thread made calls
    trows = [l for l in r.stdout.splitlines() if "@ core#" in l]
    assert len(trows) == int(args[0]), r.stdout


@pytest.mark.parametrize("ngptot,nproma,perturbation,ksplit,target_b", [(163840, 128, "", 0, 0), (120000, 64, "warm+land+conv+rain+moist+tend", 40, 3),
                                                                        (131071, 32, "sparse+warm+conv", 136, 12), (60000, 32, "hot+rain+moist+conv", 1, 1)])
def test_persistent_schedule_is_bit_identical(built, input_data, ngptot, nproma, perturbation, ksplit, target_b):
    """cloudsc_b200_set_schedule(mode 1 / mode 2 / mode 3): persistent warps that take their tiles from a queue with the input
    ring kept full across tiles (2), one launch whose last blocks are the lower halves of tiles split at a cut (3), or warps
    that run the upper and lower parts of their column tiles interleaved, the state at the cut going through a device buffer (launches big enough for the persistent path:
    >= 2 tiles per resident warp; incl. a ragged last tile; the 60 000-column case has 1.06 rounds of tiles, where only the split
    tail applies and modes 1 / 2 fall back to one CTA per tile).  Any cut level and any mix target gives the SAME bits
    (the schedule never enters the arithmetic; the CPU proof that a cut sweep equals the uncut one is
    test_cut_sweep_is_bit_identical) and, on the shipped data, the bits of the one-CTA-per-tile kernel and of the streaming
    kernel.  On synthetic inputs that reach the warm-phase code different kernels differ by rounding (4e-16: the compiler
    contracts multiply-adds per kernel), so between kernels the bar there is 1e-14 on the dwarf's error table."""
    from dwarf_p_cloudsc_b200 import params, state, api, validate
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    src = state.perturb(cols, perturbation) if perturbation else cols
    F = state.blocked_fields(src, ngptot, nproma, klev)
    api.set_params(p, k)
    st = api.State(ngptot, nproma, klev)
    outs = [f.name for f in params.OUTPUT_FIELDS]
    try:
        st.upload(F)
        assert st.get_schedule()["used"] == 3      # the default of an FP64 state with more tiles than resident warps
        res = []
        for mode, ks, tb in ((0, 0, 0), (1, ksplit, target_b), (1, max(1, klev - 1 - ksplit), 12 - min(target_b, 11)), (2, 0, 0),
                             (3, ksplit, 0), (3, max(1, klev - 1 - ksplit), 5)):
            st.set_schedule(mode, ks, tb)
            st.reset_inout()
            c0 = api.launch_counts()
            st.step(scalars["PTSPHY"])
            assert api.launch_counts()["tmap"] == c0["tmap"] + 1
            G = {n: (np.full_like(F[n], -777.0) if n in outs else F[n]) for n in F}
            st.download(G); st.sync()
            res.append(G)
    finally:
        st.close()
    for n in outs:
        assert np.array_equal(res[1][n], res[2][n]), n      # any cut, any mix: same bits (and the per-SM counters returned to zero)
        assert np.array_equal(res[4][n], res[5][n]), n      # split tail (mode 3): any cut, any number of split tiles
        if not perturbation:
            assert np.array_equal(res[1][n], res[3][n]), n  # streaming tile queue (mode 2)
            assert np.array_equal(res[1][n], res[4][n]), n  # split tail (mode 3)
            assert np.array_equal(res[0][n], res[1][n]), n
    # different kernels (one CTA per tile / persistent / streaming): same bits on the shipped data, rounding-level differences
    # where the warm-phase code runs (the compiler contracts multiply-adds per kernel)
    for other in (res[0], res[3], res[4]):
        errs = validate.validate(res[1], other, ngptot)
        assert validate.all_pass(errs) and max(e.rel_l1 for e in errs) < 1e-14, "\n" + validate.format_table(errs)


def test_c_driver_power_sampling(built):
    """EC_PMON=1 (the reference's switch, ec_pmon_mod.F90:14-57 / cloudsc_driver_mod.F90:113-118,197-199): the C driver
    reports energy and sampled power of the GPU over a sustained run of the sweep (NVML, dlopen'ed), and still validates."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "dwarf_p_cloudsc_b200", "lib", "dwarf-cloudsc-b200")
    env = dict(os.environ, EC_PMON="1")
    r = subprocess.run([exe, "1", "65536", "128"], cwd=os.path.join(root, "tests", "golden"), capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    m = re.search(r"Power usage \(sampled\):: max: (\d+) avg: ([\d.]+) count: (\d+)", r.stdout)
    assert m, r.stdout
    per = re.search(r"Power usage gpu#0 :: energy: ([\d.]+) J over (\d+) sweeps in ([\d.]+) s", r.stdout)
    assert per and int(per.group(2)) >= 16 and float(per.group(3)) >= 0.9, r.stdout
    if float(m.group(1)) > 0:          # NVML present on the box: a B200 under load draws well over 100 W
        assert 50.0 < float(m.group(2)) < 1500.0 and float(per.group(1)) > 10.0, r.stdout
    rows = [l for l in r.stdout.splitlines() if " 2D" in l]
    assert len(rows) == 21 and all(l.rstrip().endswith("PASS") for l in rows), r.stdout
    # without the switch: no power lines
    r = subprocess.run([exe, "1", "1000", "100"], cwd=os.path.join(root, "tests", "golden"), capture_output=True, text=True, timeout=600,
                       env={k: v for k, v in os.environ.items() if k != "EC_PMON"})
    assert r.returncode == 0 and "Power usage" not in r.stdout


@pytest.mark.parametrize("flags", [0, 1, 2])
def test_library_host_buffers(built, oracle, input_data, flags):
    """cloudsc_b200_host_alloc_ex: plain page-locked memory, write-combined (upload side), and 2 MB pages first-touched by
    the caller + cudaHostRegister -- run_host on arrays living in each kind gives the oracle's answer; host_free takes all."""
    from dwarf_p_cloudsc_b200 import api, params, state
    cols, scalars, klon, klev = input_data
    p, k = params.params_from_scalars(scalars)
    ngptot, nproma = 1500, 128
    F = state.blocked_fields(cols, ngptot, nproma, klev)
    R = {n: a.copy() for n, a in F.items()}
    oracle.run(R, p, k, ngptot, nproma, klev, scalars["PTSPHY"], 0)
    api.set_params(p, k)
    bufs, H = [], {}
    try:
        for f in params.FIELDS:
            if f.name not in F:
                continue
            # write-combined memory is for buffers the CPU only writes: inputs only
            hb = api.HostBuffer(F[f.name].shape, F[f.name].dtype, flags if (flags != 1 or f.use == "R") else 0)
            hb.array[...] = F[f.name]
            bufs.append(hb)
            H[f.name] = hb.array
        api.run_host(H, ngptot, nproma, klev, scalars["PTSPHY"], chunk_blocks=3, nstreams=3)
        _compare({n: np.array(a) for n, a in H.items()}, R, ngptot)
    finally:
        H.clear()
        for hb in bufs:
            hb.free()


def test_integer_pipe_compares_are_exact(built):
    """The product kernels evaluate max(x,0), min(x,1), min(x,0.5), max(x,160) and the max(x,-700) in front of every pow as
    tests on the HIGH WORD of x (one ISETP instead of a DSETP on the FP64 pipe: cloudsc_column.cuh mx0_fast / mn_hi / mx_hi /
    exp_fast_lo).  That is exact for every non-NaN x because the constants have a zero low word; checked here bit for bit
    (value for value: zeros of either sign compare equal) on zeros, denormals, infinities, the neighbours of every
    threshold and 200 000 random bit patterns.  The same call pins exp / log / reciprocal of the fast mode against
    numpy: 2e-15 relative (exp, 1/x, a/x), 1e-15 (sqrt) and 3e-16 absolute-or-relative (log)."""
    import torch
    from dwarf_p_cloudsc_b200 import api
    rng = np.random.default_rng(7)
    special = [0.0, -0.0, 5e-324, -5e-324, 2.2250738585072014e-308, -2.2250738585072014e-308, np.inf, -np.inf, 1e308, -1e308]
    for c in (0.0, 0.5, 1.0, 160.0, -700.0, 700.0):
        x = c
        for _ in range(3):
            special += [x]
            x = np.nextafter(x, np.inf)
        x = c
        for _ in range(3):
            x = np.nextafter(x, -np.inf)
            special += [x]
    bits = rng.integers(0, 2**64, size=200000, dtype=np.uint64)
    rnd = bits.view(np.float64)
    rnd = rnd[np.isfinite(rnd)]
    moderate = np.concatenate([rng.uniform(-800, 800, 50000), rng.uniform(-2, 2, 50000), 10.0 ** rng.uniform(-300, 300, 20000)])
    x = np.concatenate([np.array(special), rnd, moderate])
    out = api.selftest_primitives(torch.from_numpy(x).cuda()).cpu().numpy()
    with np.errstate(all="ignore"):
        assert np.array_equal(out[:, 0], np.where(x > 0.0, x, 0.0))
        assert np.array_equal(out[:, 1], np.where(x < 1.0, x, 1.0))
        assert np.array_equal(out[:, 2], np.where(x < 0.5, x, 0.5))
        assert np.array_equal(out[:, 3], np.where(x > 160.0, x, 160.0))
        xe = np.clip(x, -700.0, 700.0)
        ref = np.exp(xe)
        assert np.all(np.abs(out[:, 4] - ref) <= 2e-15 * ref), "exp_fast_lo: the clamp at -700 or exp itself"
        assert np.all(np.abs(out[:, 5] - ref) <= 2e-15 * ref)
        nz = (x != 0.0) & (np.abs(x) >= 2.2250738585072014e-308) & np.isfinite(x)
        lref = np.log(np.abs(x[nz]))
        assert np.all(np.abs(out[nz, 6] - lref) <= 3e-16 * np.maximum(1.0, np.abs(lref)))
        inv = nz & (np.abs(x) > 1e-300) & (np.abs(x) < 1e300)
        assert np.all(np.abs(out[inv, 7] * x[inv] - 1.0) <= 2e-15)
        sref = np.sqrt(np.abs(x[inv]))
        assert np.all(np.abs(out[inv, 8] - sref) <= 1e-15 * sref), "sqrt_fast: one cubic step on the MUFU.RSQ64H seed"
        assert np.all(np.abs(out[inv, 9] * x[inv] / 1.2345678901234567 - 1.0) <= 2e-15), "div_fast: quotient refined directly"
        # x^y = exp(y log x): |y ln x| * 3e-16 + 1e-15 relative; x = 0 only has to come out tiny (log 0 ~ -709, the exponent of
        # the exp is clamped at -700)
        ax = np.abs(x)
        for col, yy in ((10, 0.666), (11, 2.47)):
            pref = ax ** yy
            ok = np.isfinite(pref) & (pref > 1e-290) & (pref < 1e290) & (ax >= 2.2250738585072014e-308)
            bar = (np.abs(yy * np.log(ax[ok])) * 3e-16 + 1e-15) * pref[ok]
            assert np.all(np.abs(out[ok, col] - pref[ok]) <= bar), (col, yy)
            assert np.all(out[ax == 0.0, col] <= 1e-200), (col, yy)
