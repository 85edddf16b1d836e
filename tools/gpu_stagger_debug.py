"""Mode 3 (staggered launch) taken apart: which of its three kernels overlap ($CLOUDSC_B200_STAGGER_PARTS, bit 0 = A first pieces,
bit 1 = B whole tiles, bit 2 = C remainders; $CLOUDSC_B200_STAGGER_PRIO=0: equal stream priorities).
usage: python tools/gpu_stagger_debug.py [ngptot]"""
import os, sys, subprocess
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    from dwarf_p_cloudsc_b200 import api, params, state
    ngptot = int(sys.argv[2])
    cols, scalars, _, klev = state.load_input()
    p, k = params.params_from_scalars(scalars)
    api.set_params(p, k)
    pts = float(scalars["PTSPHY"])
    F = state.blocked_fields(cols, ngptot, 128, klev)
    st = api.State(ngptot, 128, klev)
    st.upload(F)

    def t(mode, nslots):
        st.set_schedule(mode, 0, nslots)
        for _ in range(3):
            st.reset_inout(); st.step(pts)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(15)]
        for a, b in ev:
            st.reset_inout(); a.record(); st.step(pts); b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in ev)
        return sum(ts) / len(ts), ts[0]
    print("   mode 0: %.4f ms (min %.4f)" % t(0, 0), flush=True)
    for nslots in (12, 4, 2):
        for parts in (7, 1, 2, 3, 6, 4):
            os.environ["CLOUDSC_B200_STAGGER_PARTS"] = str(parts)
            print("   mode 3 nslots %2d parts %d: %.4f ms (min %.4f)" % ((nslots, parts) + t(3, nslots)), flush=True)
    st.close()
    sys.exit(0)
ngptot = sys.argv[1] if len(sys.argv) > 1 else "163840"
for prio in ("1", "0"):
    print(f"stream priorities {'on' if prio == '1' else 'off'} (parts: 1 = A pieces, 2 = B whole tiles, 4 = C remainders):", flush=True)
    subprocess.run([sys.executable, __file__, "--child", ngptot], env=dict(os.environ, CLOUDSC_B200_STAGGER_PRIO=prio))
