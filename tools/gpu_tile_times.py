"""When does every 32-column tile of a launch start and end, and on which SM?  (debug build: tools/build_alt.sh tiletimes
-DCLOUDSC_TILE_TIMES, then CLOUDSC_B200_LIB=dwarf_p_cloudsc_b200/lib_alt/tiletimes/libcloudsc_b200.so python tools/gpu_tile_times.py)
Prints the distribution of tile durations, the shape of the end of the launch, and what list scheduling of the SAME durations
would give for the hardware's order, for longest-first order and for an ideal fluid machine."""
import ctypes, heapq, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dwarf_p_cloudsc_b200 import api, params, state
from dwarf_p_cloudsc_b200.params import FIELD_INDEX
cols, scalars, _, klev = state.load_input()
p, k = params.params_from_scalars(scalars)
api.set_params(p, k)
ngptot, nproma = (int(sys.argv[1]) if len(sys.argv) > 1 else 163840), 128
F = state.blocked_fields(cols, ngptot, nproma, klev)
st = api.State(ngptot, nproma, klev)
st.upload(F)
ntiles = (ngptot + 31) // 32
dbg = torch.zeros(3 * ntiles, dtype=torch.int64, device="cuda")
dev = ctypes.cast(st._h.value + 32, ctypes.POINTER(ctypes.c_void_p * 56)).contents
dev[FIELD_INDEX["TENDENCY_CML_T"]] = dbg.data_ptr()
for it in range(4):
    st.reset_inout(); dbg.zero_()
    st.step(float(scalars["PTSPHY"]))
    torch.cuda.synchronize()
t = dbg.cpu().numpy().reshape(ntiles, 3)
t0, t1, sm = t[:, 0] - t[:, 0].min(), t[:, 1] - t[:, 0].min(), t[:, 2]
dur = (t1 - t0) / 1e3
span = t1.max() / 1e3
nsm = len(np.unique(sm)); slots = 12 * nsm
print(f"{ntiles} tiles on {nsm} SMs; launch span {span:.1f} us; sum of tile durations / {slots} slots = {dur.sum() / slots:.1f} us")
print("tile duration [us]: min %.1f  p10 %.1f  median %.1f  p90 %.1f  max %.1f" % (dur.min(), *np.percentile(dur, [10, 50, 90]), dur.max()))
typ = np.arange(ntiles) % 25
print("mean duration by tile type (tile index mod 25; the 100-column pattern repeats every 25 tiles):")
print("  " + " ".join("%.0f" % dur[typ == i].mean() for i in range(25)))
# wave by wave: tiles sorted by start time
order = np.argsort(t0)
for w in range((ntiles + slots - 1) // slots):
    idx = order[w * slots:(w + 1) * slots]
    print("  wave %d: %4d tiles, start %.1f..%.1f us, end %.1f..%.1f us, mean duration %.1f us" % (w, len(idx), t0[idx].min() / 1e3, t0[idx].max() / 1e3, t1[idx].min() / 1e3, t1[idx].max() / 1e3, dur[idx].mean()))
# resident tiles over the last 300 us
for back in (300, 250, 200, 150, 100, 50, 25):
    x = span - back
    print("  %3d us before the end: %4d tiles resident (of %d slots)" % (back, int(((t0 / 1e3 <= x) & (t1 / 1e3 > x)).sum()), slots))

def list_schedule(durs):
    h = [0.0] * slots
    heapq.heapify(h)
    for d in durs:
        heapq.heappush(h, heapq.heappop(h) + d)
    return max(h)
print("list scheduling of the measured durations on %d slots: launch order %.1f us, longest first %.1f us, shortest first %.1f us, fluid %.1f us" % (
    slots, list_schedule(dur[np.argsort(t0)]), list_schedule(np.sort(dur)[::-1]), list_schedule(np.sort(dur)), dur.sum() / slots))
# per SM: mean duration of its first-wave tiles (all start at t = 0 with the same work) -- is the spread a property of the SM?
first = order[:slots]
by_sm = {}
for i in first:
    by_sm.setdefault(int(sm[i]), []).append(dur[i])
m = {s: float(np.mean(v)) for s, v in by_sm.items()}
ids = sorted(m)
print("first-wave mean tile duration per SM [us], SM id ascending (16 per row):")
for r in range(0, len(ids), 16):
    print("  %3d: " % ids[r] + " ".join("%3.0f" % m[s] for s in ids[r:r + 16]))
spread_in = np.mean([np.std(v) for v in by_sm.values()])
print("  spread between SMs: min %.0f max %.0f std %.1f us; mean spread inside an SM: %.1f us" % (min(m.values()), max(m.values()), np.std(list(m.values())), spread_in))
np.savez_compressed(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "tile_times.npz"), t0=t0, t1=t1, sm=sm)
