"""Diagnostic: which kernels agree bit for bit on the same columns?  python tools/gpu_bitwise.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dwarf_p_cloudsc_b200 import api, params, state

cols, scalars, _, klev = state.load_input()
p, k = params.params_from_scalars(scalars)
pts = float(scalars["PTSPHY"])
api.set_params(p, k)
ngptot, nproma = 1024, 128
F = state.blocked_fields(cols, ngptot, nproma, klev)

def run(variant, F=F):
    D = {n: torch.from_numpy(a).cuda() for n, a in F.items()}
    api.set_variant(variant)
    api.launch(D, ngptot, nproma, klev, pts)
    torch.cuda.synchronize()
    api.set_variant(0)
    return {n: t.cpu().numpy() for n, t in D.items()}

def run_state():
    H = {n: a.copy() for n, a in F.items()}
    st = api.State(ngptot, nproma, klev); st.upload(H); st.step(pts); st.download(H); st.sync(); st.close()
    return H

def diff(a, b, tag):
    bad = []
    for f in params.OUTPUT_FIELDS:
        x, y = a[f.name], b[f.name]
        if not np.array_equal(x, y):
            d = np.abs(x - y); i = np.unravel_index(np.argmax(d), d.shape)
            bad.append((f.name, int((x != y).sum()), float(d.max()), float(np.abs(x).max()), i))
    print(tag, "identical" if not bad else bad, flush=True)

G7, G2, G5, G9, GS = run(7), run(2), run(5), run(9), run_state()
diff(G7, GS, "tile C1 vs tmap state:")
diff(G7, G2, "tile C1 vs direct:")
diff(G7, G5, "tile C1 vs tile C4:")
diff(G7, G9, "tile C1 vs pipe:")
diff(G2, G9, "direct vs pipe:")
# permutation on the tile kernel only (1024 columns: no tail)
rng = np.random.default_rng(7); perm = rng.permutation(100)
c2 = {n: np.ascontiguousarray(a[..., perm]) for n, a in cols.items()}
P7 = run(7, state.blocked_fields(c2, ngptot, nproma, klev))
idx = perm[np.arange(ngptot) % 100]
bad = [f.name for f in params.OUTPUT_FIELDS if not np.array_equal(state.unblock(P7[f.name], ngptot), state.unblock(G7[f.name], ngptot)[..., :100][..., idx])]
print("tile C1 permutation independence:", "ok" if not bad else bad, flush=True)
