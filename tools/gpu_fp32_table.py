"""FP32 device path against the three references: per field  GPU-FP32 vs CPU-FP32 reference (oracle/_ref32 = the unmodified
kernel re-typed to float),  GPU-FP32 vs FP64 reference,  CPU-FP32 reference vs FP64 reference (the inherent FP32 error of
the scheme).  L1-relative errors, worst over the shipped data and three synthetic input sets."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import pyoracle
from dwarf_p_cloudsc_b200 import api, params, state, validate
cols, scalars, _, klev = state.load_input()
p, k = params.params_from_scalars(scalars)
pts = float(scalars["PTSPHY"])
api.set_params(p, k)
r64, r32 = pyoracle.RefOracle(), pyoracle.Ref32Oracle()
ngptot, nproma = 2000, 128
worst = {n: [0.0, 0.0, 0.0] for n in state.VALIDATE_ORDER}
for variant in ("", "warm+land+conv+rain+moist+tend", "hot+rain+moist+conv", "sparse+warm+conv"):
    src = state.perturb(cols, variant) if variant else cols
    F = state.blocked_fields(src, ngptot, nproma, klev)
    R64 = {n: a.copy() for n, a in F.items()}
    r64.run(R64, p, k, ngptot, nproma, klev, pts, 0)
    F32 = {n: (a.astype(np.float32) if a.dtype == np.float64 else a.copy()) for n, a in F.items()}
    R32 = {n: a.copy() for n, a in F32.items()}
    r32.run(R32, p, k, ngptot, nproma, klev, pts, 0)
    G = {n: a.copy() for n, a in F32.items()}
    st = api.State(ngptot, nproma, klev, api.FP32)
    st.upload(G); st.step(pts); st.download(G); st.sync(); st.close()
    print(f"--- {variant or 'shipped'}: {'field':18s} {'GPU32 vs CPU32':>15s} {'GPU32 vs FP64':>15s} {'CPU32 vs FP64':>15s}")
    for n in state.VALIDATE_ORDER:
        u = lambda A: state.unblock(np.asarray(A[n], dtype=np.float64), ngptot)
        e = [validate.field_error(n, u(G), u(R32), ngptot).rel_l1, validate.field_error(n, u(G), u(R64), ngptot).rel_l1,
             validate.field_error(n, u(R32), u(R64), ngptot).rel_l1]
        for i in range(3):
            worst[n][i] = max(worst[n][i], e[i])
        print(f"    {n:18s} {e[0]:15.3e} {e[1]:15.3e} {e[2]:15.3e}")
print("--- worst over the four input sets")
for n in state.VALIDATE_ORDER:
    print(f"    {n:18s} {worst[n][0]:15.3e} {worst[n][1]:15.3e} {worst[n][2]:15.3e}")
