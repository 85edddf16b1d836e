"""Input/reference state handling for tests and benchmarks (NumPy side).

Restates what ``load_state`` / ``load_reference`` / ``expand_*`` do in the reference
(src/cloudsc_c/cloudsc/load_state.c:69-184,279-692,701-760): read the 100-column data set, replicate
column ``g -> g mod 100`` to NGPTOT columns and block it as ``[nblocks][...][nproma]`` with NPROMA
fastest.  Padding columns of the last block are filled too (load_state.c:84-92 is commented out).
The product-side loader is the C ``load_state`` in host/load_state.c; this module exists so the
Python tests and bench.py can build the same arrays without a C driver in the loop.
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import numpy as np

from . import h5lite
from .params import FIELDS, FIELD_INDEX, params_from_scalars

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(_ROOT, "tests", "golden")

# reference.h5 dataset name -> field-table name
REFERENCE_FIELDS = {
    "PLUDE": "PLUDE", "PCOVPTOT": "PCOVPTOT", "PRAINFRAC_TOPRFZ": "PRAINFRAC_TOPRFZ",
    "PFSQLF": "PFSQLF", "PFSQIF": "PFSQIF", "PFCQLNG": "PFCQLNG", "PFCQNNG": "PFCQNNG",
    "PFSQRF": "PFSQRF", "PFSQSF": "PFSQSF", "PFCQRNG": "PFCQRNG", "PFCQSNG": "PFCQSNG",
    "PFSQLTUR": "PFSQLTUR", "PFSQITUR": "PFSQITUR", "PFPLSL": "PFPLSL", "PFPLSN": "PFPLSN",
    "PFHPSL": "PFHPSL", "PFHPSN": "PFHPSN", "TENDENCY_LOC_A": "TENDENCY_LOC_A",
    "TENDENCY_LOC_Q": "TENDENCY_LOC_Q", "TENDENCY_LOC_T": "TENDENCY_LOC_T",
    "TENDENCY_LOC_CLD": "TENDENCY_LOC_CLD",
}
# order in which cloudsc_validate prints them (cloudsc_validate.c:196-216)
VALIDATE_ORDER = ["PLUDE", "PCOVPTOT", "PRAINFRAC_TOPRFZ", "PFSQLF", "PFSQIF", "PFCQLNG", "PFCQNNG",
                  "PFSQRF", "PFSQSF", "PFCQRNG", "PFCQSNG", "PFSQLTUR", "PFSQITUR", "PFPLSL", "PFPLSN",
                  "PFHPSL", "PFHPSN", "TENDENCY_LOC_A", "TENDENCY_LOC_Q", "TENDENCY_LOC_T",
                  "TENDENCY_LOC_CLD"]


def load_input(path: str | None = None):
    """-> (columns: {field: ndarray[..., 100]}, scalars: {name: value}, klon, klev)."""
    ds = h5lite.read_h5(path or os.path.join(GOLDEN_DIR, "input.h5"))
    scalars = {k: v.reshape(-1)[0].item() for k, v in ds.items() if v.size == 1}
    klon, klev = int(scalars["KLON"]), int(scalars["KLEV"])
    cols = {f.name: ds[f.name] for f in FIELDS if f.name in ds}
    return cols, scalars, klon, klev


def load_reference(path: str | None = None) -> Dict[str, np.ndarray]:
    ds = h5lite.read_h5(path or os.path.join(GOLDEN_DIR, "reference.h5"))
    return {REFERENCE_FIELDS[k]: v for k, v in ds.items() if k in REFERENCE_FIELDS}


def nblocks_of(ngptot: int, nproma: int) -> int:
    return ngptot // nproma + min(ngptot % nproma, 1)  # cloudsc_driver.c:98


def expand(src: np.ndarray, ngptot: int, nproma: int, first_column: int = 0) -> np.ndarray:
    """Replicate ``src[..., nlon]`` to ``[nblocks, ..., nproma]`` (expand_1d/2d/3d, load_state.c:69-184).

    ``first_column`` shifts the global column index (rank offset for multi-GPU sharding,
    src/common/module/expand_mod.F90:30-46)."""
    nlon = src.shape[-1]
    nb = nblocks_of(ngptot, nproma)
    g = (first_column + np.arange(nb * nproma, dtype=np.int64)) % nlon
    out = np.take(src, g, axis=-1)                         # [..., nb*nproma]
    out = out.reshape(src.shape[:-1] + (nb, nproma))
    return np.ascontiguousarray(np.moveaxis(out, -2, 0))   # [nb, ..., nproma]


AEROSOL_FIELDS = ("PRE_ICE", "PICRIT_AER", "PNICE")   # read with yrecldp.laericesed / laericeauto


def blocked_fields(cols: Dict[str, np.ndarray], ngptot: int, nproma: int, klev: int,
                   first_column: int = 0, dtype=np.float64, live_only: bool = True,
                   aerosol: bool = False) -> Dict[str, np.ndarray]:
    """Fields in driver layout: inputs expanded, outputs zero-initialised.  With ``live_only`` the
    13 fields the kernel never touches (use N / A in the field table) are left out (NULL pointers);
    ``aerosol`` keeps the three aerosol inputs the laericesed / laericeauto branches read."""
    nb = nblocks_of(ngptot, nproma)
    out = {}
    for f in FIELDS:
        if live_only and f.use in "NA" and not (aerosol and f.name in AEROSOL_FIELDS):
            continue
        shape = (nb,) + f.level_shape(klev) + (nproma,)
        if f.name in cols:
            a = expand(cols[f.name], ngptot, nproma, first_column)
            out[f.name] = a.astype(np.int32 if f.kind == "I" else dtype, copy=False)
            assert out[f.name].shape == shape, (f.name, out[f.name].shape, shape)
        else:
            out[f.name] = np.zeros(shape, dtype=np.int32 if f.kind == "I" else dtype)
    return out


def unblock(a: np.ndarray, ngptot: int) -> np.ndarray:
    """[nb, ..., nproma] -> [..., ngptot] (valid columns only)."""
    nb, nproma = a.shape[0], a.shape[-1]
    x = np.moveaxis(a, 0, -2).reshape(a.shape[1:-1] + (nb * nproma,))
    return x[..., :ngptot]


def setup(ngptot: int, nproma: int, path: str | None = None, first_column: int = 0, dtype=np.float64):
    """Convenience: -> (fields dict, TECLDP, CloudscConsts, ptsphy, klev)."""
    cols, scalars, _klon, klev = load_input(path)
    p, k = params_from_scalars(scalars)
    return blocked_fields(cols, ngptot, nproma, klev, first_column, dtype), p, k, float(scalars["PTSPHY"]), klev


def perturb(cols: Dict[str, np.ndarray], variant: str, seed: int = 1234) -> Dict[str, np.ndarray]:
    """Synthetic variations of the 100 shipped columns that reach the branches the data set leaves
    cold (warm phase T >= 273.16 K, land points, snow detrainment, downdraughts,
    pre-existing rain).  Deterministic in (variant, seed).  Returns a new dict; inputs untouched."""
    rng = np.random.default_rng(seed)
    c = {k: v.copy() for k, v in cols.items()}
    nlon = c["PT"].shape[-1]
    klev = c["PT"].shape[0]
    lev = np.arange(klev)[:, None] / float(klev - 1)          # 0 top .. 1 surface
    parts = set(variant.split("+")) if variant else set()
    if "warm" in parts:        # melting layer, rain, Abel-Boutle evaporation, rain freezing aloft
        c["PT"] = c["PT"] + (8.0 + 14.0 * rng.random(nlon))[None, :] * lev ** 2 + 3.0
    if "hot" in parts:         # whole lower troposphere above freezing
        c["PT"] = c["PT"] + 25.0 * lev
    if "land" in parts:
        c["PLSM"] = (np.arange(nlon) % 2).astype(np.float64)
        c["KTYPE"] = ((np.arange(nlon) // 2) % 4).astype(np.int32)
    if "conv" in parts:        # snow detrainment + downdraught mass flux + more detrained water
        act = (c["PLUDE"] > 0) | (rng.random(c["PLUDE"].shape) < 0.05)
        c["PSNDE"] = np.where(act, 2.0e-6 * rng.random(c["PLUDE"].shape), 0.0)
        c["PMFD"] = -0.3 * c["PMFU"] * rng.random(c["PMFU"].shape)
        c["PLUDE"] = c["PLUDE"] * (1.0 + rng.random(c["PLUDE"].shape))
    if "rain" in parts:        # pre-existing rain / snow / supersaturation
        low = lev > 0.5
        c["PCLV"][2] = np.where(low & (rng.random(c["PT"].shape) < 0.3), 3.0e-5 * rng.random(c["PT"].shape), 0.0)
        c["PCLV"][3] = c["PCLV"][3] + np.where(rng.random(c["PT"].shape) < 0.2, 1.0e-5 * rng.random(c["PT"].shape), 0.0)
        c["PSUPSAT"] = np.where(rng.random(c["PT"].shape) < 0.1, 1.0e-5 * rng.random(c["PT"].shape), c["PSUPSAT"])
    if "moist" in parts:       # more condensation / new-cloud generation, noisy thresholds
        c["PQ"] = c["PQ"] * (1.0 + 0.35 * rng.random(c["PQ"].shape))
        c["PCLV"][0] = c["PCLV"][0] * (1.0 + rng.random(c["PT"].shape)) + np.where(
            rng.random(c["PT"].shape) < 0.15, 4.0e-5 * rng.random(c["PT"].shape), 0.0)
        c["PA"] = np.clip(c["PA"] + np.where(rng.random(c["PA"].shape) < 0.2, rng.random(c["PA"].shape), 0.0), 0.0, 1.0)
    if "tend" in parts:        # non-trivial incoming tendencies, tiny/negative condensate
        c["TENDENCY_TMP_T"] = c["TENDENCY_TMP_T"] + 2.0e-4 * (rng.random(c["PT"].shape) - 0.5)
        c["TENDENCY_TMP_CLD"][:4] = c["TENDENCY_TMP_CLD"][:4] + 2.0e-9 * (rng.random((4,) + c["PT"].shape) - 0.5)
    if "sparse" in parts:      # isolated events in otherwise clear, dry air; precipitation falling into clear layers
        shape = c["PT"].shape
        clear = rng.random((1, nlon)) < 0.5                                   # half of the columns: no cloud at all ...
        band = (lev > 0.35) & (lev < 0.55)                                    # ... except one thin layer of ice and snow aloft
        c["PA"] = np.where(clear, np.where(band, 0.3, 0.0), c["PA"])
        for m in range(4):
            c["PCLV"][m] = np.where(clear, 0.0, c["PCLV"][m])
        c["PCLV"][1] = np.where(clear & band, 2.0e-5, c["PCLV"][1])
        c["PCLV"][3] = np.where(clear & band, 1.0e-5, c["PCLV"][3])
        c["PQ"] = np.where(clear & ~band, c["PQ"] * 0.3, c["PQ"])             # dry: outside every humidity window
        c["TENDENCY_TMP_A"] = np.where(clear, 0.0, c["TENDENCY_TMP_A"])
        c["TENDENCY_TMP_CLD"][:4] = np.where(clear[None], 0.0, c["TENDENCY_TMP_CLD"][:4])
        lone = rng.random(shape) < 0.01                                       # lone sources in the clear air
        c["PSUPSAT"] = np.where(clear & lone, 3.0e-6, c["PSUPSAT"])
        lone = rng.random(shape) < 0.01
        c["PSNDE"] = np.where(clear & lone, 1.0e-6, c["PSNDE"])
        lone = rng.random(shape) < 0.01
        c["PLUDE"] = np.where(clear & lone, 5.0e-6, c["PLUDE"])
        c["PLU"] = np.where(clear & lone, 1.0e-4, c["PLU"])
        lone = rng.random(shape) < 0.01                                       # and lone moist points that open the new-cloud window
        c["PQ"] = np.where(clear & lone & (lev > 0.5), c["PQ"] * 4.0, c["PQ"])
    if "aer" in parts:         # aerosol inputs of the laericesed / laericeauto branches (identically zero in the data set)
        shape = c["PT"].shape
        c["PRE_ICE"] = 20.0 + 80.0 * rng.random(shape)                        # ice effective radius [um] -> fall speed 0.04..0.2
        c["PNICE"] = 0.027 * (0.2 + 3.0 * rng.random(shape))                  # around yrecldp.rnice
        c["PICRIT_AER"] = 3.0e-5 * (0.3 + 2.0 * rng.random(shape))            # around yrecldp.rlcritsnow
    unknown = parts - {"warm", "hot", "land", "conv", "rain", "moist", "tend", "sparse", "aer", ""}
    if unknown:
        raise ValueError(f"unknown perturbation {unknown}")
    return c
