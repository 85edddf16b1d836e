"""CPU tests of the C host side (host/load_state.c, cloudsc_validate.c, h5lite.c) through ctypes:
the loader must build exactly the arrays the reference loader builds (restated in NumPy by
dwarf_p_cloudsc_b200.state), the validator must reproduce the dwarf's table semantics."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
_dp = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def hostlib(built):
    os.environ["CLOUDSC_INPUT_H5"] = os.path.join(GOLDEN, "input.h5")
    os.environ["CLOUDSC_REFERENCE_H5"] = os.path.join(GOLDEN, "reference.h5")
    from dwarf_p_cloudsc_b200 import api  # noqa: F401  (loads libcloudsc_b200.so first: it owns the globals)
    lib = ctypes.CDLL(os.path.join(ROOT, "dwarf_p_cloudsc_b200", "lib", "libcloudsc_host.so"))
    lib.h5lite_open.restype = ctypes.c_void_p
    lib.h5lite_open.argtypes = [ctypes.c_char_p]
    lib.h5lite_close.argtypes = [ctypes.c_void_p]
    lib.h5lite_count.restype = ctypes.c_size_t
    lib.h5lite_count.argtypes = [ctypes.c_void_p]
    lib.h5lite_name.restype = ctypes.c_char_p
    lib.h5lite_name.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    lib.h5lite_read_double.argtypes = [ctypes.c_void_p, ctypes.c_char_p, _dp, ctypes.c_size_t]
    lib.h5lite_read_int.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.c_size_t]
    lib.h5lite_error.restype = ctypes.c_char_p
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_c_h5lite_matches_python_h5lite(hostlib):
    from dwarf_p_cloudsc_b200 import h5lite
    for fn in ("input.h5", "reference.h5"):
        path = os.path.join(GOLDEN, fn)
        py = h5lite.read_h5(path)
        f = hostlib.h5lite_open(path.encode())
        assert f, hostlib.h5lite_error()
        names = {hostlib.h5lite_name(f, i).decode() for i in range(hostlib.h5lite_count(f))}
        assert names == set(py)
        for name, a in py.items():
            out = np.empty(a.size, dtype=np.float64)
            assert hostlib.h5lite_read_double(f, name.encode(), out.ctypes.data_as(_dp), a.size) == 0, name
            assert np.array_equal(out, a.reshape(-1).astype(np.float64)), name
        iout = np.empty(100, dtype=np.int32)
        if "KTYPE" in py:
            assert hostlib.h5lite_read_int(f, b"/KTYPE", iout.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), 100) == 0
            assert np.array_equal(iout, py["KTYPE"])
        # errors are reported, not crashed on
        assert hostlib.h5lite_read_double(f, b"NO_SUCH_DATASET", out.ctypes.data_as(_dp), 1) == -1
        assert hostlib.h5lite_read_double(f, b"PLUDE", out.ctypes.data_as(_dp), 3) == -1
        hostlib.h5lite_close(f)
    assert not hostlib.h5lite_open(b"/nonexistent/file.h5")
    assert not hostlib.h5lite_open(os.path.join(ROOT, "bench.py").encode())   # not an HDF5 file


@pytest.mark.parametrize("ngptot,nproma,first", [(100, 100, 0), (257, 32, 0), (1000, 128, 0), (300, 64, 1234567)])
def test_c_load_state_matches_numpy_restatement(hostlib, ngptot, nproma, first):
    from dwarf_p_cloudsc_b200 import api, params, state
    cols, scalars, klon, klev = state.load_input()
    kl, kv = ctypes.c_int(), ctypes.c_int()
    hostlib.query_state(ctypes.byref(kl), ctypes.byref(kv))
    assert (kl.value, kv.value) == (klon, klev)

    want = state.blocked_fields(cols, ngptot, nproma, klev, first_column=first, live_only=False)
    got = {n: np.full_like(a, -7) for n, a in want.items()}
    p = params.TECLDP()
    ctypes.c_void_p.in_dll(api.lib, "yrecldp").value = ctypes.addressof(p)
    pts = ctypes.c_double()
    order = ["PLCRIT_AER", "PICRIT_AER", "PRE_ICE", "PCCN", "PNICE", "PT", "PQ", "TENDENCY_CML_T", "TENDENCY_CML_Q",
             "TENDENCY_CML_A", "TENDENCY_CML_CLD", "TENDENCY_TMP_T", "TENDENCY_TMP_Q", "TENDENCY_TMP_A", "TENDENCY_TMP_CLD",
             "PVFA", "PVFL", "PVFI", "PDYNA", "PDYNL", "PDYNI", "PHRSW", "PHRLW", "PVERVEL", "PAP", "PAPH", "PLSM", "KTYPE",
             "PLU", "PLUDE", "PSNDE", "PMFU", "PMFD", "PA", "PCLV", "PSUPSAT"]      # load_state.h:24-32
    hostlib.load_state_set_first_column(ctypes.c_longlong(first))
    try:
        hostlib.load_state(klon, klev, 5, ngptot, nproma, ctypes.byref(pts), *[_ptr(got[n]) for n in order])
    finally:
        hostlib.load_state_set_first_column(ctypes.c_longlong(0))
    assert pts.value == scalars["PTSPHY"]
    for n in order:
        assert np.array_equal(got[n], want[n]), n
    # parameters: *yrecldp and the YOMCST/YOETHF globals
    pw, kw = params.params_from_scalars(scalars)
    assert bytes(p) == bytes(pw)
    for n in params.KERNEL_CONSTS:
        assert ctypes.c_double.in_dll(api.lib, n).value == getattr(kw, n), n
    ctypes.c_void_p.in_dll(api.lib, "yrecldp").value = None


def test_c_validate_reference_against_itself_and_detects_errors(hostlib, capfd):
    from dwarf_p_cloudsc_b200 import state
    ngptot, nproma, klev = 257, 32, 137
    ref = state.load_reference()
    order = ["PLUDE", "PCOVPTOT", "PRAINFRAC_TOPRFZ", "PFSQLF", "PFSQIF", "PFCQLNG", "PFCQNNG", "PFSQRF", "PFSQSF",
             "PFCQRNG", "PFCQSNG", "PFSQLTUR", "PFSQITUR", "PFPLSL", "PFPLSN", "PFHPSL", "PFHPSN", "TENDENCY_LOC_A",
             "TENDENCY_LOC_Q", "TENDENCY_LOC_T", "TENDENCY_LOC_CLD"]               # cloudsc_validate.h:16-20
    F = {n: state.expand(ref[n], ngptot, nproma) for n in order}
    hostlib.cloudsc_validate(100, klev, 5, ngptot, nproma, *[_ptr(F[n]) for n in order])
    out = capfd.readouterr().out
    assert hostlib.cloudsc_validate_failures() == 0
    lines = [l for l in out.splitlines() if "2D" in l]
    assert len(lines) == 21 and all("2D1" in l and "PASS" in l for l in lines)      # iopt 1: zerrsum < eps
    assert lines[0].split()[0] == "PLUDE" and lines[-1].split()[0] == "TENDENCY_LOC%CLD"
    # one wrong value in a valid column is found; the same error in a padding column is not looked at
    F["PFHPSN"][0, 100, 3] += 1.0e-3
    F["TENDENCY_LOC_T"][-1, 50, nproma - 1] += 1.0      # block 8 holds columns 256..287, only 256 is valid
    hostlib.cloudsc_validate(100, klev, 5, ngptot, nproma, *[_ptr(F[n]) for n in order])
    out = capfd.readouterr().out
    assert hostlib.cloudsc_validate_failures() == 1
    bad = [l for l in out.splitlines() if "FAIL" in l]
    assert len(bad) == 1 and bad[0].split()[0] == "PFHPSN" and "2D3" in bad[0]      # |ref| up to 84: integer abs() sum > 0


def test_driver_executable_fails_loudly_without_a_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    exe = os.path.join(ROOT, "dwarf_p_cloudsc_b200", "lib", "dwarf-cloudsc-b200")
    r = subprocess.run([exe, "1", "100", "16"], cwd=GOLDEN, capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    r = subprocess.run([exe, "1", "2"], cwd=GOLDEN, capture_output=True, text=True)
    assert r.returncode != 0 and "right number of arguments" in r.stdout


def test_c_and_python_h5lite_agree_on_random_files(hostlib, tmp_path):
    """Property test (hypothesis): files written by the Python subset writer with a random number of datasets
    (1..90: one or several symbol-table nodes), random ranks/shapes and the four element types of input.h5 /
    reference.h5 read back identically through the Python reader and the C reader of the host driver."""
    from hypothesis import given, settings, strategies as st, HealthCheck
    from dwarf_p_cloudsc_b200 import h5lite

    dtypes = [np.float64, np.float32, np.int32, np.int64]
    shapes = st.lists(st.integers(1, 6), min_size=1, max_size=3).map(tuple)
    counter = [0]

    @settings(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck))
    @given(n=st.integers(1, 90), seed=st.integers(0, 2 ** 31 - 1), shape_list=st.lists(shapes, min_size=1, max_size=8),
           dt_list=st.lists(st.sampled_from(range(4)), min_size=1, max_size=8))
    def check(n, seed, shape_list, dt_list):
        rng = np.random.default_rng(seed)
        ds = {}
        for i in range(n):
            shp = shape_list[i % len(shape_list)]
            dt = dtypes[dt_list[i % len(dt_list)]]
            a = (rng.standard_normal(shp) * 1e3) if np.issubdtype(dt, np.floating) else rng.integers(-2 ** 30, 2 ** 30, shp)
            ds[f"{'FIELD' if i % 3 else 'x'}_{i:03d}"] = np.ascontiguousarray(a.astype(dt))
        counter[0] += 1
        path = str(tmp_path / f"r{counter[0]}.h5")
        h5lite.write_h5(path, ds)
        back = h5lite.read_h5(path)
        assert set(back) == set(ds)
        f = hostlib.h5lite_open(path.encode())
        assert f, hostlib.h5lite_error()
        try:
            assert hostlib.h5lite_count(f) == n
            for name, a in ds.items():
                assert back[name].dtype == a.dtype and np.array_equal(back[name], a), name
                out = np.empty(a.size, dtype=np.float64)
                assert hostlib.h5lite_read_double(f, name.encode(), out.ctypes.data_as(_dp), a.size) == 0, (name, hostlib.h5lite_error())
                assert np.array_equal(out, a.reshape(-1).astype(np.float64)), name
        finally:
            hostlib.h5lite_close(f)
        os.remove(path)

    check()


def test_validate_replicated_equals_the_plain_validator():
    """bench.py validates every column of every rank's shard against the replicated reference without building a
    full-size copy of it (validate.validate_replicated); same numbers as the plain table on an explicit expansion,
    incl. rank offsets, ragged last blocks and NGPTOT below one replication period."""
    import numpy as np
    from dwarf_p_cloudsc_b200 import state, validate
    ref = state.load_reference()
    rng = np.random.default_rng(0)
    for ngptot, nproma, fc in ((1000, 128, 0), (7000, 64, 300), (3333, 32, 17), (250, 100, 5), (90, 16, 0)):
        R = {n: state.expand(a, ngptot, nproma, fc) for n, a in ref.items()}
        G = {n: a + 1e-9 * rng.standard_normal(a.shape) * np.abs(a).max() for n, a in R.items()}
        for a, b in zip(validate.validate(G, R, ngptot), validate.validate_replicated(G, ref, ngptot, nproma, fc)):
            assert a.name == b.name
            for x, y in zip(a[1:], b[1:]):
                assert x == y or abs(x - y) <= 1e-12 * max(abs(x), abs(y)), (ngptot, nproma, a.name, a, b)


def test_c_validate_fails_on_nan_in_an_all_zero_field(hostlib, capfd):
    """fmin/fmax drop NaN, and a field whose reference is identically zero (PFPLSL, PFHPSL in the data set) is judged
    on max|err| alone: a NaN there must still FAIL."""
    from dwarf_p_cloudsc_b200 import state
    ngptot, nproma, klev = 100, 50, 137
    ref = state.load_reference()
    order = ["PLUDE", "PCOVPTOT", "PRAINFRAC_TOPRFZ", "PFSQLF", "PFSQIF", "PFCQLNG", "PFCQNNG", "PFSQRF", "PFSQSF",
             "PFCQRNG", "PFCQSNG", "PFSQLTUR", "PFSQITUR", "PFPLSL", "PFPLSN", "PFHPSL", "PFHPSN", "TENDENCY_LOC_A",
             "TENDENCY_LOC_Q", "TENDENCY_LOC_T", "TENDENCY_LOC_CLD"]
    F = {n: state.expand(ref[n], ngptot, nproma) for n in order}
    assert not F["PFPLSL"].any()
    F["PFPLSL"][1, 7, 3] = np.nan
    F["PFHPSL"][...] = np.nan
    hostlib.cloudsc_validate(100, klev, 5, ngptot, nproma, *[_ptr(F[n]) for n in order])
    cap = capfd.readouterr()
    assert hostlib.cloudsc_validate_failures() == 2
    bad = sorted(l.split()[0] for l in cap.out.splitlines() if "FAIL" in l)
    assert bad == ["PFHPSL", "PFPLSL"], cap.out
    assert "non-finite" in cap.err


def test_c_h5lite_rejects_a_truncated_or_inconsistent_file(hostlib, tmp_path):
    """Element counts come from the dataspace, the bytes that exist from the layout message: a file cut short, or one
    whose dataspace claims more elements than its storage holds, is refused instead of read out of bounds."""
    from dwarf_p_cloudsc_b200 import h5lite
    rng = np.random.default_rng(3)
    a = rng.random((7, 100))
    good = tmp_path / "good.h5"
    h5lite.write_h5(str(good), {"A": a, "N": np.array([5], dtype=np.int32)})
    raw = bytearray(good.read_bytes())
    out = np.empty(a.size)
    f = hostlib.h5lite_open(str(good).encode())
    assert f and hostlib.h5lite_read_double(f, b"A", out.ctypes.data_as(_dp), a.size) == 0 and np.array_equal(out, a.reshape(-1))
    hostlib.h5lite_close(f)
    # (1) cut the file in the middle of A's data
    cut = tmp_path / "cut.h5"
    pos = bytes(raw).find(a.tobytes())
    assert pos > 0
    cut.write_bytes(bytes(raw[:pos + a.nbytes // 2]))
    f = hostlib.h5lite_open(str(cut).encode())
    if f:
        assert hostlib.h5lite_read_double(f, b"A", out.ctypes.data_as(_dp), a.size) == -1
        hostlib.h5lite_close(f)
    # (2) dataspace of A claims 7 x 200 elements, its contiguous storage still holds 7 x 100
    import struct
    bad = bytearray(raw)
    dims = struct.pack("<QQ", 7, 100)
    i = bytes(bad).find(dims)
    assert i > 0
    bad[i:i + 16] = struct.pack("<QQ", 7, 200)
    lie = tmp_path / "lie.h5"
    lie.write_bytes(bytes(bad))
    f = hostlib.h5lite_open(str(lie).encode())
    assert f
    big = np.empty(1400)
    assert hostlib.h5lite_read_double(f, b"A", big.ctypes.data_as(_dp), 1400) == -1
    assert b"exceeds its storage" in hostlib.h5lite_error()
    hostlib.h5lite_close(f)


@pytest.mark.parametrize("args", [("2", "100", "16"), ("1", "300", "64")])
def test_reference_executable_runs_on_our_loader(built, args):
    """oracle/_ref/dwarf-cloudsc-c = the reference's OWN dwarf-cloudsc-c (dwarf_cloudsc.c, cloudsc_driver.c, cloudsc_c.c,
    cloudsc_validate.c, yo*_c.c, all unmodified) with host/load_state.c + host/h5lite.c standing in for load_state.c +
    libhdf5.  Its own error table against reference.h5 must be clean: this pins the loader end to end (input.h5 ->
    expand_* -> the reference kernel -> load_reference) and is the executable behind BASELINE configs[0]."""
    exe = os.path.join(ROOT, "oracle", "_ref", "dwarf-cloudsc-c")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dwarf-cloudsc-c not built (needs /root/reference or a prebuilt copy)")
    env = dict(os.environ, OMP_NUM_THREADS=args[0], OMP_STACKSIZE="256M")
    r = subprocess.run([exe, *args], cwd=GOLDEN, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert f"NUMOMP={args[0]}, NGPTOT={args[1]}, NPROMA={args[2]}" in r.stdout
    rows = [l.split() for l in r.stdout.splitlines() if len(l.split()) >= 7 and l.split()[1][:2] == "2D"]
    assert len(rows) == 21, r.stdout
    for row in rows:
        vmin, vmax, maxerr, avgerr, relerr = map(float, row[2:7])
        assert np.isfinite([vmin, vmax, maxerr, avgerr, relerr]).all(), row
        assert maxerr <= 1e-12 * max(abs(vmin), abs(vmax), 1e-300) or maxerr == 0.0, row     # C kernel vs the Fortran-made file: <= 3e-13
    assert "TOTAL" in r.stdout
