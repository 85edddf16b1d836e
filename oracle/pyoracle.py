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
    flags = "-O2 -ffp-contract=off"


def _cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


class RefFastOracle(_Oracle):
    """The same unmodified reference kernel, built for speed (-O3, -march=x86-64-v4 where the host has AVX-512, else
    -v3; FP contraction on).  TIMING ONLY: parity is always judged against RefOracle (-O2, contraction off)."""
    blocks_fn = "cloudsc_ref_blocks"
    kind = "reference"

    def __init__(self):
        flags = _cpu_flags()
        v4 = {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= flags
        v3 = {"avx2", "fma", "bmi2"} <= flags
        if not v3:
            raise OSError("host CPU lacks AVX2/FMA: no fast reference build for it")
        self.level = "v4" if v4 else "v3"
        self.libname = os.path.join("_ref", f"libcloudsc_ref_O3{self.level}.so")
        self.flags = f"-O3 -march=x86-64-{self.level} -ffp-contract=fast"
        super().__init__()


class Ref32Oracle:
    """oracle/_ref/libcloudsc_ref32.so: the unmodified reference kernel re-typed to single precision by the preprocessor
    (oracle/ref32_wrap.c) -- what a `--single-precision` build of the C dwarf would be.  ``fields`` are float32 arrays
    (KTYPE int32) in the driver layout; parameters are narrowed from the double-precision objects."""
    kind = "reference-fp32"

    def __init__(self):
        path = os.path.join(HERE, "_ref", "libcloudsc_ref32.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        os.environ.setdefault("OMP_STACKSIZE", "1G")
        self.lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        self.lib.cloudsc_ref32_blocks.restype = ctypes.c_double
        self.lib.cloudsc_ref32_blocks.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                                  ctypes.POINTER(ctypes.c_void_p)]
        self.lib.cloudsc_ref32_set_params.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                                      ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                                      ctypes.POINTER(ctypes.c_double)]

    def run(self, fields, p, k, ngptot, nproma, klev, ptsphy, nthreads=0):
        from dwarf_p_cloudsc_b200.params import FIELDS, KERNEL_CONSTS, TECLDP_INTS, TECLDP_REALS
        assert self.lib.cloudsc_ref32_sizeof_tecldp() == 4 * (len(TECLDP_REALS) + len(TECLDP_INTS) + 3), "re-typed struct TECLDP layout"
        reals = (ctypes.c_double * len(TECLDP_REALS))(*[getattr(p, n) for n in TECLDP_REALS])
        ints = (ctypes.c_int * len(TECLDP_INTS))(*[getattr(p, n) for n in TECLDP_INTS])
        kk = (ctypes.c_double * len(KERNEL_CONSTS))(*[getattr(k, n) for n in KERNEL_CONSTS])
        self.lib.cloudsc_ref32_set_params(reals, len(TECLDP_REALS), ints, len(TECLDP_INTS), p.nshapep, p.nshapeq, p.nbeta, kk)
        tab = (ctypes.c_void_p * len(FIELDS))()
        for f in FIELDS:
            a = fields.get(f.name)
            if a is None:
                continue
            assert a.flags["C_CONTIGUOUS"] and a.dtype == (np.int32 if f.kind == "I" else np.float32), f.name
            tab[f.index] = a.ctypes.data
        t = self.lib.cloudsc_ref32_blocks(int(nthreads), int(ngptot), int(nproma), int(klev), float(ptsphy), tab)
        if t < 0:
            raise ValueError("oracle rejected the arguments")
        return t


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


class GpuRef:
    """The reference's OWN CUDA kernels (src/cloudsc_cuda/cloudsc/cloudsc_c_k_caching.cu / cloudsc_c.cu, unmodified,
    compiled for sm_100a by `make -C oracle gpuref`) behind oracle/gpu_ref_harness.cu: the same-device GPU baseline.
    ``fields`` are FP64 device arrays (torch tensors or raw pointers) in the driver layout, KLEV = 137."""
    libs = {"kcaching": "libcloudsc_gpuref_kcaching.so", "scc": "libcloudsc_gpuref_scc.so"}
    sources = {"kcaching": "src/cloudsc_cuda/cloudsc/cloudsc_c_k_caching.cu", "scc": "src/cloudsc_cuda/cloudsc/cloudsc_c.cu"}

    def __init__(self, which="kcaching"):
        path = os.path.join(HERE, "_ref", self.libs[which])
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.which = which
        self.source = self.sources[which]
        self.lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        self.lib.cloudsc_gpuref_run.restype = ctypes.c_int
        self.lib.cloudsc_gpuref_run.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.POINTER(ctypes.c_double),
                                                ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_int,
                                                ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
        self.lib.cloudsc_gpuref_kernel_info.argtypes = [ctypes.POINTER(ctypes.c_int)]

    def kernel_info(self):
        out = (ctypes.c_int * 4)()
        rc = self.lib.cloudsc_gpuref_kernel_info(out)
        if rc:
            raise RuntimeError(f"cloudsc_gpuref_kernel_info: CUDA error {rc}")
        return {"registers": out[0], "local_bytes": out[1], "static_smem": out[2], "max_threads_per_block": out[3]}

    def run(self, fields, p, k, ngptot, nproma, ptsphy, plude0=None, warmup=0, reps=1):
        """-> {"ms": mean, "ms_min": ..., "ms_max": ...} over ``reps`` launches; outputs are left in ``fields``."""
        from dwarf_p_cloudsc_b200.params import FIELDS, KERNEL_CONSTS
        assert ctypes.sizeof(p) == self.lib.cloudsc_gpuref_tecldp_bytes(), "struct TECLDP layout differs"
        tab = (ctypes.c_void_p * len(FIELDS))()
        for f in FIELDS:
            a = fields.get(f.name)
            if a is not None:
                tab[f.index] = a.data_ptr() if hasattr(a, "data_ptr") else int(a)
        kk = (ctypes.c_double * 28)(*[getattr(k, n) for n in KERNEL_CONSTS])
        ms = (ctypes.c_float * 3)()
        p0 = None if plude0 is None else ctypes.c_void_p(plude0.data_ptr() if hasattr(plude0, "data_ptr") else int(plude0))
        rc = self.lib.cloudsc_gpuref_run(tab, ctypes.addressof(p), kk, int(ngptot), int(nproma), float(ptsphy), p0,
                                         int(warmup), int(reps), ms)
        if rc:
            raise RuntimeError(f"cloudsc_gpuref_run: CUDA error {rc}")
        return {"ms": ms[0], "ms_min": ms[1], "ms_max": ms[2]}
