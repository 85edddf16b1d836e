"""Where a warp's time goes, level by level (debug build: tools/build_alt.sh timing -DCLOUDSC_TIMING, then
CLOUDSC_B200_LIB=dwarf_p_cloudsc_b200/lib_alt/timing/libcloudsc_b200.so python tools/gpu_level_timing.py).
The tensor-map kernel stamps clock64() at: level begin, after the mbarrier wait, before and after the refill issue."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dwarf_p_cloudsc_b200 import api, params, state
cols, scalars, _, klev = state.load_input()
p, k = params.params_from_scalars(scalars)
api.set_params(p, k)
ngptot, nproma = (int(sys.argv[1]) if len(sys.argv) > 1 else 163840), 128
F = state.blocked_fields(cols, ngptot, nproma, klev)
st = api.State(ngptot, nproma, klev)
st.upload(F)
dbg = torch.zeros(10 * 4 * klev, dtype=torch.int64, device="cuda")
# the state never allocates TENDENCY_CML_T (use N): point its table entry at the debug buffer
from dwarf_p_cloudsc_b200.params import FIELD_INDEX
tab_ptr = ctypes.cast(st._h, ctypes.POINTER(ctypes.c_void_p))
# struct cloudsc_b200_state { int x7; void *dev[56]; ... }: dev[] starts at byte 32
dev = ctypes.cast(st._h.value + 32, ctypes.POINTER(ctypes.c_void_p * 56)).contents
dev[FIELD_INDEX["TENDENCY_CML_T"]] = dbg.data_ptr()
for it in range(3):
    st.reset_inout(); dbg.zero_()
    st.step(float(scalars["PTSPHY"]))
    torch.cuda.synchronize()
t = dbg.cpu().numpy().reshape(10, klev, 4)
def seg(x, lo, hi): return x[lo:hi].mean()
for c in range(min(10, (ngptot // 32 + 511) // 512)):
    beg, aft, endin, endout = (t[c, :, i] for i in range(4))
    lev = np.diff(beg); wait = aft - beg; body = endin - aft; issue = endout - endin
    print("CTA %4d total %8d | level: top14 %6.0f k14-80 %6.0f k81+ %6.0f | wait %5.0f %5.0f %5.0f | body %6.0f %6.0f %6.0f | refill issue %5.0f %5.0f %5.0f" % (
        c * 512, beg[-1] - beg[0], seg(lev, 0, 14), seg(lev, 14, 80), seg(lev, 81, 136),
        seg(wait, 0, 14), seg(wait, 14, 80), seg(wait, 81, 137), seg(body, 0, 14), seg(body, 14, 80), seg(body, 81, 137),
        seg(issue, 0, 14), seg(issue, 14, 80), seg(issue, 81, 137)))

