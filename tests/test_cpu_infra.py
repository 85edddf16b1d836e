"""CPU-side tests (`-m "not gpu"`): fixtures, HDF5-lite, ABI surface, host logic.  No compute call
on the CUDA library is made here (there is no GPU in the build container)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_h5lite_reads_reference_h5_as_shipped():
    from dwarf_p_cloudsc_b200 import h5lite
    ds = h5lite.read_h5(os.path.join(ROOT, "tests", "golden", "reference.h5"))
    assert len(ds) == 24
    assert ds["KLEV"].item() == 137 and ds["KLON"].item() == 100
    assert ds["PLUDE"].shape == (137, 100) and ds["PLUDE"].dtype == np.float64
    assert ds["PFSQLF"].shape == (138, 100)
    assert ds["TENDENCY_LOC_CLD"].shape == (5, 137, 100)
    assert ds["PRAINFRAC_TOPRFZ"].shape == (100,)
    assert np.isfinite(ds["TENDENCY_LOC_T"]).all() and np.abs(ds["PFHPSN"]).max() > 80.0


def test_h5lite_roundtrip(tmp_path):
    from dwarf_p_cloudsc_b200 import h5lite
    rng = np.random.default_rng(0)
    ds = {"A": rng.random((3, 7)), "B_int": np.arange(5, dtype=np.int32), "C": np.array([2.5]),
          "D64": np.array([137], dtype=np.int64), "E_f32": rng.random((2, 3, 4)).astype(np.float32),
          "Mixed_Case_Name": np.array([1.0])}
    for i in range(40):                      # more than one symbol-table node
        ds[f"S{i:03d}"] = np.array([float(i)])
    path = str(tmp_path / "t.h5")
    h5lite.write_h5(path, ds)
    back = h5lite.read_h5(path)
    assert set(back) == set(ds)
    for k in ds:
        assert back[k].dtype == ds[k].dtype and np.array_equal(back[k], ds[k]), k


def test_input_fixture_schema():
    """input.h5 follows serialbox2hdf5.py:35-56: 40 fields (C order) + scalars as shape-(1,) datasets."""
    from dwarf_p_cloudsc_b200 import state, params
    cols, scalars, klon, klev = state.load_input()
    assert (klon, klev) == (100, 137)
    assert cols["PT"].shape == (137, 100) and cols["PAPH"].shape == (138, 100)
    assert cols["PCLV"].shape == (5, 137, 100) and cols["KTYPE"].dtype == np.int32
    assert scalars["PTSPHY"] == 3600.0 and scalars["YRECLDP_NCLDTOP"] == 15 and scalars["YRECLDP_NSSOPT"] == 1
    for f in params.INPUT_FIELDS:
        assert f.name in cols, f.name
    p, k = params.params_from_scalars(scalars)
    assert p.ncldtop == 15 and abs(k.rtt - 273.16) < 1e-12 and p.rcl_kkaac > 0
    # value ranges quoted in SURVEY 8d
    assert 196.0 < cols["PT"].min() and cols["PT"].max() < 268.0 and (cols["PLSM"] == 0).all()


def test_params_layout_matches_reference_abi():
    from dwarf_p_cloudsc_b200 import params
    assert ctypes.sizeof(params.TECLDP) == 912              # SURVEY 2.1 #3 [probe]
    assert params.TECLDP.lcldextra.offset == 816 and params.TECLDP.nshapep.offset == 888
    assert len(params.TECLDP_REALS) == 102 and len(params.TECLDP_INTS) == 17
    assert params.NFIELDS == 56
    names = [f.name for f in params.FIELDS]
    assert names[0] == "PT" and names[26] == "KTYPE" and names[28] == "PLUDE" and names[55] == "PFHPSN"
    # the swapped pair in the cloudsc_c() argument list (cloudsc_c.h:27-29; SURVEY App. A note)
    assert names.index("PFCQNNG") < names.index("PFCQLNG")
    assert params.algorithmic_bytes_per_column(137) == 56036
    assert params.algorithmic_bytes_per_column(137, 4) == 28020 - 0  # FP32: 28 020 B


def test_expand_replicates_like_load_state():
    """expand_2d: field[b][l][i] = buffer[l][(b*nproma+i) % nlon], padding columns filled too
    (load_state.c:123-138)."""
    from dwarf_p_cloudsc_b200 import state
    src = np.arange(3 * 7, dtype=np.float64).reshape(3, 7)
    out = state.expand(src, ngptot=18, nproma=4)
    assert out.shape == (5, 3, 4)
    for b in range(5):
        for l in range(3):
            for i in range(4):
                assert out[b, l, i] == src[l, (b * 4 + i) % 7]
    out2 = state.expand(src, ngptot=8, nproma=4, first_column=5)
    assert out2[0, 1, 0] == src[1, 5] and out2[1, 2, 3] == src[2, (5 + 7) % 7]
    u = state.unblock(out, 18)
    assert u.shape == (3, 18) and np.array_equal(u[:, 7:14], src)


def test_validate_table_semantics():
    """The reference's MaxRelErr-% uses integer abs() on doubles (cloudsc_validate.c:74)."""
    from dwarf_p_cloudsc_b200 import validate
    ref = np.array([0.5, 0.25, 0.75, 0.1])
    got = ref + np.array([0.0, 1e-6, 0.0, 0.0])
    e = validate.field_error("X", got, ref, 4)
    assert e.iopt == 2 and abs(e.relerr_c - 100 * 1e-6) < 1e-12 and e.flagged     # zsum truncated to 0
    assert abs(e.rel_l1 - 1e-6 / 1.6) < 1e-12 and not e.passes()
    e0 = validate.field_error("X", ref, ref, 4)
    assert e0.iopt == 1 and e0.relerr_c == 0 and e0.passes()
    big = np.array([10.0, 20.0])
    e3 = validate.field_error("X", big * (1 + 1e-12), big, 2)
    assert e3.iopt == 3 and e3.passes()
    assert not validate.field_error("X", np.array([np.nan, 1.0]), big, 2).passes()


def test_abi_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "cloudsc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cloudsc_b200_\w+|cloudsc_c)\s*\(", hdr))
    declared.discard("cloudsc_b200_state")
    assert len(declared) >= 33 and {"cloudsc_b200_run_host_ex", "cloudsc_b200_release_thread", "cloudsc_b200_launch_counts"} <= declared, sorted(declared)
    lib = ctypes.CDLL(os.path.join(ROOT, "dwarf_p_cloudsc_b200", "lib", "libcloudsc_b200.so"))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/cloudsc_b200.h but not exported"
    # the reference-style globals of the drop-in contract
    for g in ("yrecldp", "rg", "rtt", "r2es", "rkoop2", "nclv"):
        ctypes.c_int.in_dll(lib, g)


def test_abi_field_table_and_errors_without_gpu(built):
    from dwarf_p_cloudsc_b200 import api, params
    for f in params.FIELDS:
        assert api.lib.cloudsc_b200_field_name(f.index).decode() == f.name
        assert api.lib.cloudsc_b200_field_kind(f.index).decode() == f.kind
        assert api.lib.cloudsc_b200_field_use(f.index).decode() == f.use
        assert api.lib.cloudsc_b200_field_block_elems(f.index, 128, 137) == f.block_elems(128, 137)
    # argument errors are reported, not crashed on
    with pytest.raises(api.CloudscError) as ei:
        api.launch({}, 0, 128, 137, 3600.0)
    assert ei.value.code == -2
    with pytest.raises(api.CloudscError):
        api.launch({}, 100, 100, 137, 3600.0)      # NULL fields
    cols_p = params.TECLDP()
    cols_p.ncldtop = 1
    with pytest.raises(api.CloudscError) as ei:
        api.set_params(cols_p, params.CloudscConsts())
    assert ei.value.code == -1


def test_no_product_code_touches_the_oracle():
    """oracle/ is test infrastructure: the package must not import, link or execute it."""
    pkg = os.path.join(ROOT, "dwarf_p_cloudsc_b200")
    for d, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".c", ".h", ".cu", ".cuh")):
                txt = open(os.path.join(d, fn)).read()
                assert "pyoracle" not in txt and "libcloudsc_oracle" not in txt and "libcloudsc_ref" not in txt \
                    and "host_emul" not in txt.replace("tests/host_emul", ""), os.path.join(d, fn)
