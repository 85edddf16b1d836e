"""Per-level cycle counts of a few CTAs (debug build with -DCLOUDSC_TIMING, tools/build_alt.sh timing -DCLOUDSC_TIMING)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dwarf_p_cloudsc_b200 import api, params, state
cols, scalars, _, klev = state.load_input()
p, k = params.params_from_scalars(scalars)
api.set_params(p, k)
if len(sys.argv) > 1:
    api.set_variant(int(sys.argv[1]))
ngptot, nproma = 163840, 128
F = state.blocked_fields(cols, ngptot, nproma, klev)
D = {n: torch.from_numpy(a).cuda() for n, a in F.items()}
plude0 = D["PLUDE"].clone()
dbg = torch.zeros(10 * 4 * klev, dtype=torch.int64, device="cuda")
D["TENDENCY_CML_T"] = dbg
for it in range(3):
    D["PLUDE"].copy_(plude0); dbg.zero_()
    api.launch(D, ngptot, nproma, klev, float(scalars["PTSPHY"]))
    torch.cuda.synchronize()
t = dbg.cpu().numpy().reshape(10, klev, 4)
def seg(x, lo, hi): return x[lo:hi].mean()
for c in range(10):
    beg, aft, endin, endout = (t[c, :, i] for i in range(4))
    lev = np.diff(beg); wait = aft - beg; body = endin - aft; issue = endout - endin
    print("CTA %4d total %8d | level: top14 %6.0f k14-80 %6.0f k81+ %6.0f | wait %5.0f %5.0f %5.0f | body %6.0f %6.0f %6.0f | level_end(TMA issue) %5.0f %5.0f %5.0f" % (
        c * 512, beg[-1] - beg[0], seg(lev, 0, 14), seg(lev, 14, 80), seg(lev, 81, 136),
        seg(wait, 0, 14), seg(wait, 14, 80), seg(wait, 81, 137), seg(body, 0, 14), seg(body, 14, 80), seg(body, 81, 137),
        seg(issue, 0, 14), seg(issue, 14, 80), seg(issue, 81, 137)))
