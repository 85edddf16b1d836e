"""In-tree build of the native code (explicit nvcc/gcc, no JIT cache).

    python -m dwarf_p_cloudsc_b200.build            # everything that is out of date
Outputs (git-ignored, but they travel to the GPU box with the gpurun snapshot):
    dwarf_p_cloudsc_b200/lib/libcloudsc_b200.so     kernels + C ABI
    dwarf_p_cloudsc_b200/lib/dwarf-cloudsc-b200     C driver with the dwarf-cloudsc-c command line
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
LIBDIR = os.path.join(PKG, "lib")
INC = os.path.join(ROOT, "include")
LIB = os.path.join(LIBDIR, "libcloudsc_b200.so")
DRIVER = os.path.join(LIBDIR, "dwarf-cloudsc-b200")
HOSTLIB = os.path.join(LIBDIR, "libcloudsc_host.so")   # loader + validator as a library (ctypes tests)

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
HOSTCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"   # $CC here is a wrapper without libgomp
HOSTCXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-ccbin", HOSTCXX, f"-I{INC}", f"-I{CSRC}"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def _glob(d, exts):
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def build_library(force=False, verbose=False, extra_nvcc=()):
    os.makedirs(LIBDIR, exist_ok=True)
    headers = _glob(CSRC, (".h", ".cuh")) + _glob(INC, (".h",))
    cu = _glob(CSRC, (".cu",))
    objs = []
    for src in cu:
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ".o")
        if force or _newer(obj, [src] + headers):
            _run([NVCC, *ARCH, *NVFLAGS, *extra_nvcc, "-c", src, "-o", obj], verbose)
        objs.append(obj)
    gobj = os.path.join(LIBDIR, "cloudsc_globals.o")
    gsrc = os.path.join(HOST, "cloudsc_globals.c")
    if force or _newer(gobj, [gsrc] + headers):
        _run([HOSTCC, "-std=c11", "-O2", "-fPIC", f"-I{INC}", "-c", gsrc, "-o", gobj], verbose)
    objs.append(gobj)
    if force or _newer(LIB, objs):
        _run([NVCC, *ARCH, "-shared", "-ccbin", HOSTCXX, "-o", LIB, *objs, "-lcudart"], verbose)
    return LIB


def build_driver(force=False, verbose=False):
    srcs = [os.path.join(HOST, f) for f in ("dwarf_cloudsc.c", "cloudsc_driver.c", "load_state.c",
                                            "cloudsc_validate.c", "h5lite.c", "mycpu.c", "ec_pmon.c")]
    if not all(os.path.exists(s) for s in srcs):
        return None
    headers = _glob(HOST, (".h",)) + _glob(INC, (".h",))
    if force or _newer(DRIVER, srcs + headers + [LIB]):
        _run([HOSTCC, "-std=c11", "-O2", "-fopenmp", f"-I{INC}", f"-I{HOST}", *srcs, "-o", DRIVER,
              f"-L{LIBDIR}", "-lcloudsc_b200", f"-Wl,-rpath,$ORIGIN", "-lm", "-ldl"], verbose)
    return DRIVER


def build_hostlib(force=False, verbose=False):
    """load_state / load_reference / cloudsc_validate / h5lite as a shared object, so that the CPU tests
    can drive the C loader and validator through ctypes (the driver executable links the same sources)."""
    srcs = [os.path.join(HOST, f) for f in ("load_state.c", "cloudsc_validate.c", "h5lite.c", "mycpu.c")]
    headers = _glob(HOST, (".h",)) + _glob(INC, (".h",))
    if force or _newer(HOSTLIB, srcs + headers + [LIB]):
        _run([HOSTCC, "-std=c11", "-O2", "-fopenmp", "-fPIC", "-shared", f"-I{INC}", f"-I{HOST}", *srcs, "-o", HOSTLIB,
              f"-L{LIBDIR}", "-lcloudsc_b200", "-Wl,-rpath,$ORIGIN", "-lm"], verbose)
    return HOSTLIB


def build_all(force=False, verbose=False):
    lib = build_library(force, verbose)
    drv = build_driver(force, verbose)
    build_hostlib(force, verbose)
    return lib, drv


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose=True))
