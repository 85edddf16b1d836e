#!/usr/bin/env python
"""Regenerate the golden fixtures under tests/golden/ from the reference checkout.

Runs ONLY in the build container (needs /root/reference); the GPU box uses the committed output.

* tests/golden/input.h5      <- /root/reference/data/input_*.dat + MetaData-input.json, following the
                                schema of serialbox2hdf5/serialbox2hdf5.py:35-56 (fields transposed to C
                                order, every scalar a shape-(1,) dataset; bool/int as int32, which is what
                                load_state.c:27-33,671-690 reads them as).  The upstream input.h5 blob is
                                not part of the checkout (.MISSING_LARGE_BLOBS).
* tests/golden/reference.h5  <- byte copy of config-files/reference.h5 after checking that each of its 21
                                payloads is bit-identical to data/reference_<FIELD>.dat.
"""
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from dwarf_p_cloudsc_b200 import h5lite  # noqa: E402

REF = os.environ.get("CLOUDSC_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

_DT = {1: np.uint8, 2: np.dtype("<i4"), 5: np.dtype("<f8")}


def read_serialbox(prefix):
    meta = json.load(open(os.path.join(REF, "data", f"MetaData-{prefix}.json")))
    fields = {}
    for name, fm in meta["field_map"].items():
        dims = fm["dims"]
        raw = np.fromfile(os.path.join(REF, "data", f"{prefix}_{name}.dat"), dtype=_DT[fm["type_id"]])
        # Serialbox "Binary" archive = Fortran order of `dims`  ->  C order of reversed dims
        fields[name] = raw.reshape(dims[::-1])
    scalars = {k: v["value"] for k, v in meta["global_meta_info"].items()}
    return fields, scalars


def main():
    os.makedirs(OUT, exist_ok=True)
    fields, scalars = read_serialbox("input")
    ds = {}
    for k, a in fields.items():
        ds[k] = a.astype(np.int32) if a.dtype == np.uint8 else a
    for k, v in scalars.items():
        if isinstance(v, bool) or isinstance(v, int):
            ds[k] = np.array([int(v)], dtype="<i4")
        else:
            ds[k] = np.array([float(v)], dtype="<f8")
    ds.setdefault("KLON", np.array([100], dtype="<i4"))
    ds.setdefault("KLEV", np.array([137], dtype="<i4"))
    h5lite.write_h5(os.path.join(OUT, "input.h5"), ds)
    back = h5lite.read_h5(os.path.join(OUT, "input.h5"))
    assert set(back) == set(ds)
    for k in ds:
        assert back[k].dtype == ds[k].dtype and np.array_equal(back[k], ds[k]), k

    ref_fields, _ = read_serialbox("reference")
    src = os.path.join(REF, "config-files", "reference.h5")
    h5 = h5lite.read_h5(src)
    for k, a in ref_fields.items():
        assert np.array_equal(h5[k], a), f"reference.h5:{k} differs from reference_{k}.dat"
    shutil.copyfile(src, os.path.join(OUT, "reference.h5"))
    os.chmod(os.path.join(OUT, "reference.h5"), 0o644)
    print("wrote", OUT, {f: os.path.getsize(os.path.join(OUT, f)) for f in ("input.h5", "reference.h5")})


if __name__ == "__main__":
    main()
