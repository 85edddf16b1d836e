"""Persistent-kernel schedule sweep: step time of the resident state for (mode, ksplit, target_b), every setting checked
bit for bit against the one-CTA-per-tile kernel.  usage: python tools/gpu_persist_sweep.py [ngptot] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dwarf_p_cloudsc_b200 import api, params, state

_pos = [a for a in sys.argv[1:] if not a.startswith("--")]
ngptot = int(_pos[0]) if len(_pos) > 0 else 163840
steps = int(_pos[1]) if len(_pos) > 1 else 20
nproma = 128
cols, scalars, _, klev = state.load_input()
p, k = params.params_from_scalars(scalars)
api.set_params(p, k)
pts = float(scalars["PTSPHY"])
F = state.blocked_fields(cols, ngptot, nproma, klev)
st = api.State(ngptot, nproma, klev)
st.upload(F)
outs = [f.name for f in params.OUTPUT_FIELDS]


def run(mode, ks, tb):
    st.set_schedule(mode, ks, tb)
    for _ in range(3):
        st.reset_inout(); st.step(pts)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        st.reset_inout()
        a.record(); st.step(pts); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    G = {n: np.empty_like(F[n]) for n in outs}
    G.update({n: F[n] for n in F if n not in G})
    st.download(G); st.sync()
    return sum(ts) / len(ts), ts[0], {n: G[n] for n in outs}


base_ms, base_min, base = run(0, 0, 0)
print(f"mode 0 (one CTA per tile)          : {base_ms:.4f} ms (min {base_min:.4f})", flush=True)
settings = [(2, 0, 0), (1, 1, 12), (1, 136, 12), (1, 82, 12), (1, 90, 6), (2, 0, 0)]
if "--split" in sys.argv:
    settings = [(3, 82, 4), (3, 82, 3), (3, 82, 2), (3, 82, 5), (3, 95, 4), (3, 105, 4), (3, 70, 4), (3, 60, 4), (3, 95, 3), (3, 105, 6), (3, 82, 4)]
if "--confirm" in sys.argv:      # the default of an FP64 state (3, 0, 0) beside its neighbours and one CTA per tile, alternated
    settings = [(3, 0, 0), (0, 0, 0), (3, 82, 5), (3, 82, 6), (3, 82, 4), (0, 0, 0), (3, 0, 0)]
if "--small" in sys.argv:        # launches below two rounds of tiles (run with CLOUDSC_B200_SPLIT_MIN_TILES=1): where does splitting stop paying
    settings = [(3, 0, 0), (3, 82, 3), (3, 82, 8), (3, 82, 12), (3, 60, 5), (3, 100, 5), (0, 0, 0), (3, 0, 0)]
if "--full" in sys.argv:
    settings = [(1, 0, 0)] + [(1, ks, tb) for ks in (82, 70, 90, 60) for tb in (4, 6, 7, 8, 12, 1)]
for mode, ks, tb in settings:
    ms, mn, G = run(mode, ks, tb)
    same = all(np.array_equal(G[n], base[n]) for n in outs)
    print(f"mode {mode} ksplit {ks:3d} target_b {tb:2d}        : {ms:.4f} ms (min {mn:.4f})  x{base_ms / ms:.3f}  bitwise {'OK' if same else 'DIFFERENT'}", flush=True)
ms, mn, _ = run(0, 0, 0)
print(f"mode 0 again                       : {ms:.4f} ms (min {mn:.4f})")
