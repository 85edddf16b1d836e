"""Python binding of the C ABI in ``include/cloudsc_b200.h`` (ctypes over libcloudsc_b200.so).

This is the host-side mirror used by the tests, ``bench.py`` and ``__graft_entry__``; the compute
path is always the CUDA library -- there is no Python/NumPy/torch fallback, and a missing library
raises at import of this module.  torch is used by callers only for device memory, streams and
``torch.distributed``; this module takes raw pointers (``tensor.data_ptr()`` / ``ndarray.ctypes.data``).
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Mapping, Optional, Sequence

from .params import FIELDS, NFIELDS, CloudscConsts, TECLDP

_HERE = os.path.dirname(os.path.abspath(__file__))
# $CLOUDSC_B200_LIB: an alternative build of the same library (compiler-flag experiments, tools/build_alt.sh)
LIB_PATH = os.environ.get("CLOUDSC_B200_LIB") or os.path.join(_HERE, "lib", "libcloudsc_b200.so")

FP64, FP32 = 0, 1
VARIANT_AUTO = 0   # see csrc/cloudsc_kernels.cuh for the other kernel variants


class CloudscError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cloudsc_b200 error {code}: {msg}")
        self.code = code


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m dwarf_p_cloudsc_b200.build` "
            "(there is no CPU fallback for the CLOUDSC hot path)")
    return ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_LOCAL)


lib = _load()
_FT = ctypes.c_void_p * NFIELDS
_vp, _i, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_double

lib.cloudsc_b200_version.restype = ctypes.c_char_p
lib.cloudsc_b200_last_error.restype = ctypes.c_char_p
lib.cloudsc_b200_field_name.restype = ctypes.c_char_p
lib.cloudsc_b200_field_name.argtypes = [_i]
lib.cloudsc_b200_field_kind.restype = ctypes.c_char
lib.cloudsc_b200_field_kind.argtypes = [_i]
lib.cloudsc_b200_field_use.restype = ctypes.c_char
lib.cloudsc_b200_field_use.argtypes = [_i]
lib.cloudsc_b200_field_block_elems.restype = ctypes.c_size_t
lib.cloudsc_b200_field_block_elems.argtypes = [_i, _i, _i]
lib.cloudsc_b200_set_params.argtypes = [ctypes.POINTER(TECLDP), ctypes.POINTER(CloudscConsts)]
lib.cloudsc_b200_launch.argtypes = [_FT, _i, _i, _i, _d, _i, _vp]
lib.cloudsc_b200_kernel_info.argtypes = [_i, _i, ctypes.POINTER(ctypes.c_int * 8)]
lib.cloudsc_b200_run_host.argtypes = [_FT, _i, _i, _i, _d, _i, _i, ctypes.POINTER(ctypes.c_double * 4)]
lib.cloudsc_b200_run_host_ex.argtypes = [_FT, _i, _i, _i, _d, _i, _i, _i, ctypes.POINTER(ctypes.c_double * 4)]
lib.cloudsc_b200_create.argtypes = [_i, _i, _i, _i, ctypes.POINTER(_vp)]
lib.cloudsc_b200_destroy.argtypes = [_vp]
lib.cloudsc_b200_upload.argtypes = [_vp, _FT, _vp]
lib.cloudsc_b200_download.argtypes = [_vp, _FT, _vp]
lib.cloudsc_b200_step.argtypes = [_vp, _d, _i, _i, _vp]
lib.cloudsc_b200_reset_inout.argtypes = [_vp, _vp]
lib.cloudsc_b200_set_schedule.argtypes = [_vp, _i, _i, _i]
lib.cloudsc_b200_set_schedule.restype = _i
lib.cloudsc_b200_get_schedule.argtypes = [_vp, ctypes.POINTER(ctypes.c_int * 4)]
lib.cloudsc_b200_get_schedule.restype = _i
lib.cloudsc_b200_field_ptr.restype = _vp
lib.cloudsc_b200_field_ptr.argtypes = [_vp, _i]
lib.cloudsc_b200_sync.argtypes = [_vp]
lib.cloudsc_b200_set_variant.argtypes = [_i]
lib.cloudsc_b200_state_kernel_info.argtypes = [_vp, ctypes.POINTER(ctypes.c_int * 8)]
lib.cloudsc_b200_state_kernel_info.restype = _i
lib.cloudsc_c.argtypes = [_i, _i, _i, _i, _d] + [_vp] * NFIELDS
lib.cloudsc_b200_host_alloc_ex.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
lib.cloudsc_b200_host_alloc_ex.restype = _i
lib.cloudsc_b200_host_free.argtypes = [_vp]
lib.cloudsc_b200_host_free.restype = _i
lib.cloudsc_b200_selftest_primitives.argtypes = [_vp, _i, _vp, _vp]
lib.cloudsc_b200_selftest_primitives.restype = _i
lib.cloudsc_b200_launch_counts.argtypes = [ctypes.POINTER(ctypes.c_longlong * 5)]
for _n in ("set_params", "launch", "kernel_info", "launch_counts", "run_host", "run_host_ex", "release_thread", "release_cache", "create", "destroy", "upload", "download",
           "step", "reset_inout", "sync", "set_variant", "get_variant", "device_count",
           "set_params_from_globals"):
    getattr(lib, "cloudsc_b200_" + _n).restype = _i
lib.cloudsc_c.restype = _i


def _check(rc: int):
    if rc != 0:
        raise CloudscError(rc, lib.cloudsc_b200_last_error().decode())


def version() -> str:
    return lib.cloudsc_b200_version().decode()


def device_count() -> int:
    return lib.cloudsc_b200_device_count()


def _ptr(x) -> Optional[int]:
    if x is None:
        return None
    if hasattr(x, "data_ptr"):          # torch.Tensor
        return x.data_ptr()
    if hasattr(x, "ctypes"):            # numpy.ndarray
        return x.ctypes.data
    return int(x)


def field_table(fields: Mapping[str, object]) -> "_FT":
    """dict name -> tensor/ndarray/int  ->  void*[56] in cloudsc_c() argument order."""
    tab = _FT()
    for f in FIELDS:
        tab[f.index] = _ptr(fields.get(f.name))
    return tab


def set_params(p: TECLDP, k: CloudscConsts) -> None:
    """Snapshot the parameter objects into __constant__ memory of the current CUDA device."""
    _check(lib.cloudsc_b200_set_params(ctypes.byref(p), ctypes.byref(k)))


def set_variant(v: int) -> None:
    _check(lib.cloudsc_b200_set_variant(v))


def launch(dev_fields: Mapping[str, object], ngptot: int, nproma: int, klev: int, ptsphy: float,
           precision: int = FP64, stream: int = 0) -> None:
    """One kernel launch over all NPROMA blocks of device-resident fields (asynchronous)."""
    _check(lib.cloudsc_b200_launch(field_table(dev_fields), ngptot, nproma, klev, ptsphy, precision,
                                   ctypes.c_void_p(stream)))


def launch_table(tab, ngptot, nproma, klev, ptsphy, precision=FP64, stream=0) -> None:
    """Same, with a prebuilt field table (keeps ctypes marshalling out of timed loops)."""
    _check(lib.cloudsc_b200_launch(tab, ngptot, nproma, klev, ptsphy, precision, ctypes.c_void_p(stream)))


def run_host(host_fields: Mapping[str, object], ngptot: int, nproma: int, klev: int, ptsphy: float,
             chunk_blocks: int = 0, nstreams: int = 0, precision: int = FP64) -> Dict[str, float]:
    """End-to-end run on HOST arrays (pinned preferred): chunked H2D || kernel || D2H pipeline.
    precision=FP32: float host arrays, single-precision device arithmetic (cloudsc_b200_run_host_ex)."""
    stats = (ctypes.c_double * 4)()
    if precision == FP64:
        _check(lib.cloudsc_b200_run_host(field_table(host_fields), ngptot, nproma, klev, ptsphy,
                                         chunk_blocks, nstreams, ctypes.byref(stats)))
    else:
        _check(lib.cloudsc_b200_run_host_ex(field_table(host_fields), ngptot, nproma, klev, ptsphy, precision,
                                            chunk_blocks, nstreams, ctypes.byref(stats)))
    return {"h2d_bytes": stats[0], "d2h_bytes": stats[1], "kernel_ms": stats[2], "wall_ms": stats[3]}


KERNEL_FAMILIES = ("tmap", "tile", "direct", "pipe", "direct_aer")


def selftest_primitives(x):
    """The fast-mode primitives of the kernels evaluated on the device for every element of the CUDA double tensor `x`
    (cloudsc_b200_selftest_primitives): returns a [n, 12] tensor."""
    import torch
    assert x.is_cuda and x.dtype == torch.float64 and x.is_contiguous()
    out = torch.empty((x.numel(), 12), dtype=torch.float64, device=x.device)
    _check(lib.cloudsc_b200_selftest_primitives(ctypes.c_void_p(x.data_ptr()), x.numel(), ctypes.c_void_p(out.data_ptr()),
                                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


def launch_counts() -> Dict[str, int]:
    """Kernels launched by this process so far, per kernel family (cloudsc_b200_launch_counts)."""
    out = (ctypes.c_longlong * 5)()
    _check(lib.cloudsc_b200_launch_counts(out))
    return dict(zip(KERNEL_FAMILIES, (int(x) for x in out)))


HOST_WRITE_COMBINED, HOST_HUGEPAGES = 1, 2


class HostBuffer:
    """Page-locked host memory from the library (cloudsc_b200_host_alloc_ex), viewed as a NumPy array."""

    def __init__(self, shape, dtype, flags: int = 0):
        import numpy as np
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = ctypes.c_void_p()
        _check(lib.cloudsc_b200_host_alloc_ex(ctypes.byref(self._p), self.nbytes, flags))
        buf = (ctypes.c_char * max(self.nbytes, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            lib.cloudsc_b200_host_free(self._p)
            self._p = ctypes.c_void_p()


def release_cache() -> None:
    """Free the streams and staging arenas run_host keeps between calls (every device)."""
    _check(lib.cloudsc_b200_release_cache())


def release_thread() -> None:
    """Free the calling thread's cloudsc_c() staging buffers."""
    _check(lib.cloudsc_b200_release_thread())


def kernel_info(nproma: int = 128, precision: int = FP64) -> Dict[str, int]:
    out = (ctypes.c_int * 8)()
    _check(lib.cloudsc_b200_kernel_info(nproma, precision, ctypes.byref(out)))
    keys = ("registers", "local_bytes", "static_smem", "max_threads_per_block", "blocks_per_sm",
            "block_threads", "variant", "dynamic_smem")
    return dict(zip(keys, out))


def cloudsc_c(kidia: int, kfdia: int, klon: int, klev: int, ptsphy: float, host_fields: Mapping[str, object]) -> int:
    """The binary-compatible drop-in (host pointers, ONE block, parameters from the C globals)."""
    tab = field_table(host_fields)
    return lib.cloudsc_c(kidia, kfdia, klon, klev, ptsphy, *[tab[i] for i in range(NFIELDS)])


def set_globals(p: TECLDP, k: CloudscConsts) -> None:
    """Fill the library's reference-style globals (what load_state() does in the C driver)."""
    from .params import KERNEL_CONSTS
    set_globals._keep = p
    ctypes.c_void_p.in_dll(lib, "yrecldp").value = ctypes.addressof(p)
    for n in KERNEL_CONSTS:
        ctypes.c_double.in_dll(lib, n).value = getattr(k, n)
    for n, v in (("nclv", 5), ("ncldql", 1), ("ncldqi", 2), ("ncldqr", 3), ("ncldqs", 4), ("ncldqv", 5)):
        ctypes.c_int.in_dll(lib, n).value = v


class State:
    """Resident device state for a fixed geometry (``cloudsc_b200_create`` ... ``destroy``)."""

    def __init__(self, ngptot: int, nproma: int, klev: int, precision: int = FP64):
        self._h = ctypes.c_void_p()
        _check(lib.cloudsc_b200_create(ngptot, nproma, klev, precision, ctypes.byref(self._h)))
        self.ngptot, self.nproma, self.klev = ngptot, nproma, klev

    def upload(self, host_fields, stream: int = 0):
        _check(lib.cloudsc_b200_upload(self._h, field_table(host_fields), ctypes.c_void_p(stream)))

    def download(self, host_fields, stream: int = 0):
        _check(lib.cloudsc_b200_download(self._h, field_table(host_fields), ctypes.c_void_p(stream)))

    def step(self, ptsphy: float, first_block: int = 0, nblocks: int = 0, stream: int = 0):
        _check(lib.cloudsc_b200_step(self._h, ptsphy, first_block, nblocks, ctypes.c_void_p(stream)))

    def set_schedule(self, mode: int, ksplit: int = 0, target_b: int = 0):
        """0: one CTA per 32-column tile; 1: persistent warps, upper / lower parts of their tiles interleaved; 2: persistent warps, tile
        queue; 3: split tail (default of an FP64 state)."""
        _check(lib.cloudsc_b200_set_schedule(self._h, mode, ksplit, target_b))

    SCHEDULE_NAMES = {0: "one CTA per tile", 1: "persistent, interleaved pieces", 2: "persistent, tile queue", 3: "split tail"}

    def get_schedule(self) -> Dict[str, object]:
        """The schedule as set and the one a step() over the whole state actually gets (`used`: small launches, aerosol runs
        and states without tensor maps keep one CTA per tile)."""
        out = (ctypes.c_int * 4)()
        _check(lib.cloudsc_b200_get_schedule(self._h, ctypes.byref(out)))
        return {"mode": out[0], "ksplit": out[1], "target_b": out[2], "used": out[3], "used_name": self.SCHEDULE_NAMES[out[3]]}

    def reset_inout(self, stream: int = 0):
        _check(lib.cloudsc_b200_reset_inout(self._h, ctypes.c_void_p(stream)))

    def kernel_info(self) -> Dict[str, int]:
        out = (ctypes.c_int * 8)()
        _check(lib.cloudsc_b200_state_kernel_info(self._h, ctypes.byref(out)))
        keys = ("registers", "local_bytes", "static_smem", "max_threads_per_block", "blocks_per_sm",
                "block_threads", "variant", "dynamic_smem")
        return dict(zip(keys, out))

    def field_ptr(self, name: str) -> int:
        from .params import FIELD_INDEX
        return lib.cloudsc_b200_field_ptr(self._h, FIELD_INDEX[name])

    def sync(self, stream: int = 0):
        _check(lib.cloudsc_b200_sync(ctypes.c_void_p(stream)))

    def close(self):
        if self._h:
            lib.cloudsc_b200_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
