"""ctypes front-end of tests/host_emul/libcloudsc_emul.so (test-only host build of the device column code)."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Emul:
    def __init__(self, fast=False):
        """fast=False: CLOUDSC_STRICT=1 build (reference evaluation order); fast=True: the product
        kernels' arithmetic mode (shared reciprocals, polynomial exp)."""
        name = "libcloudsc_emul_fast.so" if fast else "libcloudsc_emul.so"
        self.lib = ctypes.CDLL(os.path.join(HERE, name), mode=ctypes.RTLD_LOCAL)
        for n in ("cloudsc_emul_blocks_f64", "cloudsc_emul_blocks_f32", "cloudsc_emul_blocks_parked_f64"):
            fn = getattr(self.lib, n)
            fn.restype = ctypes.c_int
            fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int,
                           ctypes.c_int, ctypes.c_int, ctypes.c_double]

        fn = self.lib.cloudsc_emul_blocks_ranged_f64
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int,
                       ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int]

        fn = self.lib.cloudsc_emul_split_tail_f64
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                       ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int * 4)]
        self.lib.cloudsc_emul_split_piece.restype = None
        self.lib.cloudsc_emul_split_piece.argtypes = [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_int * 5)]
        self.lib.cloudsc_emul_split_count.restype = ctypes.c_int
        self.lib.cloudsc_emul_split_count.argtypes = [ctypes.c_int] * 6
        self.lib.cloudsc_emul_split_cut.restype = ctypes.c_int
        self.lib.cloudsc_emul_split_cut.argtypes = [ctypes.c_int] * 2

    def _table(self, fields, dtype=np.float64):
        from dwarf_p_cloudsc_b200.params import FIELDS
        tab = (ctypes.c_void_p * len(FIELDS))()
        for f in FIELDS:
            a = fields.get(f.name)
            if a is None:
                continue
            assert a.flags["C_CONTIGUOUS"] and a.dtype == (np.int32 if f.kind == "I" else dtype), f.name
            tab[f.index] = a.ctypes.data
        return tab

    def run_split_tail(self, fields, p, k, ngptot, nproma, klev, ptsphy, nsm, per_sm, ksplit=0, target_b=0):
        """The split-tail launch (schedule mode 3) on a pretend machine of nsm SMs x per_sm resident one-warp CTAs: the block
        plan of csrc/cloudsc_split_plan.h walked in dispatch order.  Returns {blocks, nsplit, cut, carry_loads}."""
        info = (ctypes.c_int * 4)()
        rc = self.lib.cloudsc_emul_split_tail_f64(ctypes.addressof(p), ctypes.addressof(k), self._table(fields), ngptot, nproma, klev,
                                                  ptsphy, nsm, per_sm, ksplit, target_b, ctypes.byref(info))
        if rc:
            raise RuntimeError(f"split-tail emulation returned {rc}")
        return dict(zip(("blocks", "nsplit", "cut", "carry_loads"), info))

    def split_piece(self, b, ntiles, nsplit, cut, klev):
        out = (ctypes.c_int * 5)()
        self.lib.cloudsc_emul_split_piece(b, ntiles, nsplit, cut, klev, ctypes.byref(out))
        return dict(zip(("tile", "kb", "ke", "lower", "upper"), out))

    def split_count(self, ntiles, nsm, per_sm, target_b=0, max_flags=4096, carry_tiles=1 << 30):
        return self.lib.cloudsc_emul_split_count(ntiles, nsm, per_sm, target_b, max_flags, carry_tiles)

    def split_cut(self, ksplit, klev):
        return self.lib.cloudsc_emul_split_cut(ksplit, klev)

    def run(self, fields, p, k, ngptot, nproma, klev, ptsphy, dtype=np.float64, parked=False, cuts=None):
        """parked=True: the IO policy has per-thread scratch, so the column code runs the code path of the TMA kernels
        (long-lived state through IO::scr)."""
        from dwarf_p_cloudsc_b200.params import FIELDS
        tab = (ctypes.c_void_p * len(FIELDS))()
        for f in FIELDS:
            a = fields.get(f.name)
            if a is None:
                continue
            assert a.flags["C_CONTIGUOUS"] and a.dtype == (np.int32 if f.kind == "I" else dtype), f.name
            tab[f.index] = a.ctypes.data
        fn = self.lib.cloudsc_emul_blocks_f64 if dtype == np.float64 else self.lib.cloudsc_emul_blocks_f32
        if parked:
            assert dtype == np.float64
            fn = self.lib.cloudsc_emul_blocks_parked_f64
        if cuts is not None:      # the sweep of every column cut at the two given levels (the persistent kernel's code path)
            assert dtype == np.float64
            rc = self.lib.cloudsc_emul_blocks_ranged_f64(ctypes.addressof(p), ctypes.addressof(k), tab, ngptot, nproma, klev,
                                                         ptsphy, int(cuts[0]), int(cuts[1]))
        else:
            rc = fn(ctypes.addressof(p), ctypes.addressof(k), tab, ngptot, nproma, klev, ptsphy)
        if rc:
            raise RuntimeError(f"emulation returned {rc}")
