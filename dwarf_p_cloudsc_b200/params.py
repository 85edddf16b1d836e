"""ctypes mirrors of the dwarf's parameter objects and field table.

Generated at import time from the X-macro tables in ``include/cloudsc_params.h`` and
``include/cloudsc_b200.h`` (single source of truth; reference layout: yoecldp_c.h:21-141,
yomcst_c.h:14-72, yoethf_c.h:14-36, argument order cloudsc_c.h:18-29).
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, NamedTuple

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE_DIR = os.path.join(_ROOT, "include")


def _macro_body(text: str, name: str) -> str:
    m = re.search(r"#define\s+%s\(X\)\s*\\\n((?:.*\\\n)*.*)\n" % re.escape(name), text)
    if not m:
        raise RuntimeError(f"macro {name} not found")
    return m.group(1)


def _names(text: str, name: str) -> List[str]:
    return re.findall(r"X\((\w+)\)", _macro_body(text, name))


_params_h = open(os.path.join(INCLUDE_DIR, "cloudsc_params.h")).read()
_b200_h = open(os.path.join(INCLUDE_DIR, "cloudsc_b200.h")).read()

TECLDP_REALS = _names(_params_h, "CLOUDSC_TECLDP_REALS")
TECLDP_INTS = _names(_params_h, "CLOUDSC_TECLDP_INTS")
YOMCST_REALS = _names(_params_h, "CLOUDSC_YOMCST_REALS")
YOETHF_REALS = _names(_params_h, "CLOUDSC_YOETHF_REALS")
KERNEL_CONSTS = _names(_params_h, "CLOUDSC_KERNEL_CONSTS")


class TECLDP(ctypes.Structure):
    """``struct TECLDP`` (912 bytes)."""
    _fields_ = ([(n, ctypes.c_double) for n in TECLDP_REALS] +
                [(n, ctypes.c_int) for n in TECLDP_INTS] +
                [("nshapep", ctypes.c_double), ("nshapeq", ctypes.c_double), ("nbeta", ctypes.c_int)])


class CloudscConsts(ctypes.Structure):
    """``struct cloudsc_consts``: the YOMCST/YOETHF members the kernel reads."""
    _fields_ = [(n, ctypes.c_double) for n in KERNEL_CONSTS]


class Field(NamedTuple):
    index: int
    name: str
    kind: str   # L [klev], H [klev+1], C [nclv][klev], S scalar-per-column double, I per-column int
    use: str    # R read, W written, B in/out, N never touched, A aerosol-only

    def block_elems(self, nproma: int, klev: int) -> int:
        return {"L": klev, "H": klev + 1, "C": 5 * klev, "S": 1, "I": 1}[self.kind] * nproma

    def level_shape(self, klev: int):
        return {"L": (klev,), "H": (klev + 1,), "C": (5, klev), "S": (), "I": ()}[self.kind]


FIELDS: List[Field] = [Field(i, n, k, u) for i, (n, k, u) in
                       enumerate(re.findall(r"X\((\w+),\s*(\w),\s*(\w)\)", _macro_body(_b200_h, "CLOUDSC_FIELDS")))]
FIELD_INDEX: Dict[str, int] = {f.name: f.index for f in FIELDS}
NFIELDS = len(FIELDS)
assert NFIELDS == 56, NFIELDS
INPUT_FIELDS = [f for f in FIELDS if f.use in "RB"]
OUTPUT_FIELDS = [f for f in FIELDS if f.use in "WB"]

# bytes one column moves through HBM at FP64 (SURVEY 8d / BASELINE.md section 2): every R/B field read
# once, every W/B field written once.
def algorithmic_bytes_per_column(klev: int = 137, real_bytes: int = 8) -> int:
    rd = sum(f.block_elems(1, klev) * (4 if f.kind == "I" else real_bytes) for f in FIELDS if f.use in "RB")
    # the kernel reads 4 of the 5 species of PCLV / TENDENCY_TMP_CLD
    rd -= 2 * klev * real_bytes
    wr = sum(f.block_elems(1, klev) * real_bytes for f in FIELDS if f.use in "WB")
    return rd + wr


FLOP_PER_COLUMN = 12482329.0 / 100.0  # cloudsc_driver.c:101,245


def params_from_scalars(scalars: Dict[str, float]):
    """Build (TECLDP, CloudscConsts) from the scalar datasets of input.h5 (load_state.c:538-690)."""
    p = TECLDP()
    for n in TECLDP_REALS + ["nshapep", "nshapeq"]:
        setattr(p, n, float(_lookup(scalars, "YRECLDP_" + n)))
    for n in TECLDP_INTS + ["nbeta"]:
        setattr(p, n, int(_lookup(scalars, "YRECLDP_" + n)))
    k = CloudscConsts()
    for n in KERNEL_CONSTS:
        setattr(k, n, float(_lookup(scalars, n)))
    return p, k


def _lookup(scalars, name):
    # dataset names keep the Fortran spelling (e.g. YRECLDP_RCL_KKAac); match case-insensitively
    if not hasattr(_lookup, "_cache") or _lookup._cache[0] is not scalars:
        _lookup._cache = (scalars, {k.upper(): v for k, v in scalars.items()})
    return _lookup._cache[1][name.upper()]
