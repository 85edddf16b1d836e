"""Which of (nproma, perturbation, ksplit, target_b) makes schedule mode 1 differ from mode 0, and where."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dwarf_p_cloudsc_b200 import api, params, state
cols, scalars, _, klev = state.load_input()
p, k = params.params_from_scalars(scalars)
api.set_params(p, k)
pts = float(scalars["PTSPHY"])
outs = [f.name for f in params.OUTPUT_FIELDS]
for ngptot, nproma, pert in ((120000, 64, "warm+land+conv+rain+moist+tend"), (120000, 128, "warm+land+conv+rain+moist+tend"), (120000, 64, "")):
    src = state.perturb(cols, pert) if pert else cols
    F = state.blocked_fields(src, ngptot, nproma, klev)
    st = api.State(ngptot, nproma, klev)
    st.upload(F)
    def run(mode, ks, tb):
        st.set_schedule(mode, ks, tb); st.reset_inout(); st.step(pts)
        G = {n: (np.full_like(F[n], -777.0) if n in outs else F[n]) for n in F}
        st.download(G); st.sync()
        return G
    base = run(0, 0, 0)
    base2 = run(0, 0, 0)
    print(ngptot, nproma, pert or "shipped", "mode0 reproducible:", all(np.array_equal(base[n], base2[n]) for n in outs), "finite:", all(np.isfinite(base[n]).all() for n in outs), flush=True)
    for ks, tb in ((40, 3), (40, 12), (82, 3), (82, 12), (41, 12), (13, 12), (14, 12), (20, 12)):
        G = run(1, ks, tb)
        bad = [n for n in outs if not np.array_equal(G[n], base[n])]
        msg = "OK"
        if bad:
            n = bad[0]
            d = np.argwhere(G[n] != base[n])
            lev = sorted(set(int(x[-2]) for x in d)) if G[n].ndim >= 3 else []
            cols_ = sorted(set(int(x[0]) * nproma + int(x[-1]) for x in d))
            msg = f"{len(bad)} fields differ; {n}: {len(d)} points, levels {lev[:6]}..{lev[-3:]}, columns {cols_[:8]} (n={len(cols_)}), max rel {np.nanmax(np.abs(G[n]-base[n]))/np.abs(base[n]).max():.2e}"
        print(f"   ksplit {ks:3d} target {tb:2d}: {msg}", flush=True)
    st.close()
