"""ctypes front-end of the CPU checkers -- TEST INFRASTRUCTURE ONLY.

May be imported by tests/, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of bench.py, never by the product package.

* ``RefOracle``   wraps oracle/_ref/libcloudsc_ref.so  = the UNMODIFIED reference kernel
                  (src/cloudsc_c/cloudsc/cloudsc_c.c compiled by oracle/Makefile).
* ``PortOracle``  wraps oracle/libcloudsc_oracle.so    = our C restatement (oracle/cloudsc_oracle.c).
Both expose ``run(fields, p, k, ngptot, nproma, klev, ptsphy, nthreads) -> seconds`` operating IN PLACE
on a dict of blocked NumPy arrays (see dwarf_p_cloudsc_b200.state.blocked_fields).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def build(targets=("port",), quiet=True):
    subprocess.run(["make", "-C", HERE, *targets], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class _Oracle:
    libname = None
    blocks_fn = None

    def __init__(self):
        path = os.path.join(HERE, self.libname)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        os.environ.setdefault("OMP_STACKSIZE", "1G")  # ~51 KB of VLAs per column
        self.lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        self.fn = getattr(self.lib, self.blocks_fn)
        self.fn.restype = ctypes.c_double
        self.fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                            ctypes.POINTER(ctypes.c_void_p)]
        self._keep = None

    def set_params(self, p, k):
        """Fill the kernel's globals (yrecldp, rg, r2es, ...) the way load_state.c:538-690 does."""
        from dwarf_p_cloudsc_b200.params import KERNEL_CONSTS
        self._keep = p
        ctypes.c_void_p.in_dll(self.lib, "yrecldp").value = ctypes.addressof(p)
        for n in KERNEL_CONSTS:
            ctypes.c_double.in_dll(self.lib, n).value = getattr(k, n)
        for n, v in (("nclv", 5), ("ncldql", 1), ("ncldqi", 2), ("ncldqr", 3), ("ncldqs", 4), ("ncldqv", 5)):
            ctypes.c_int.in_dll(self.lib, n).value = v

    def run(self, fields, p, k, ngptot, nproma, klev, ptsphy, nthreads=0):
        from dwarf_p_cloudsc_b200.params import FIELDS
        self.set_params(p, k)
        tab = (ctypes.c_void_p * len(FIELDS))()
        for f in FIELDS:
            a = fields.get(f.name)
            if a is None:
                continue
            assert a.flags["C_CONTIGUOUS"] and a.dtype == (np.int32 if f.kind == "I" else np.float64), f.name
            tab[f.index] = a.ctypes.data
        t = self.fn(int(nthreads), int(ngptot), int(nproma), int(klev), float(ptsphy), tab)
        if t < 0:
            raise ValueError("oracle rejected the arguments")
        return t


class RefOracle(_Oracle):
    libname = os.path.join("_ref", "libcloudsc_ref.so")
    blocks_fn = "cloudsc_ref_blocks"
    kind = "reference"


class PortOracle(_Oracle):
    libname = "libcloudsc_oracle.so"
    blocks_fn = "cloudsc_oracle_blocks"
    kind = "port"


def best_oracle():
    """The reference build when it travelled with the snapshot, else the port."""
    try:
        return RefOracle()
    except (FileNotFoundError, OSError):
        return PortOracle()
