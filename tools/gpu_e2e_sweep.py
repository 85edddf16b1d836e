"""run_host pipeline tuning: wall time of the end-to-end call for several (chunk_blocks, nstreams)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dwarf_p_cloudsc_b200 import api, params, state
cols, scalars, _, klev = state.load_input()
p, k = params.params_from_scalars(scalars)
api.set_params(p, k)
ngptot, nproma = 163840, 128
pts = float(scalars["PTSPHY"])
F = state.blocked_fields(cols, ngptot, nproma, klev)
H = {n: torch.from_numpy(a).pin_memory() for n, a in F.items()}
plude0 = H["PLUDE"].clone()
MODES = [("2", "0"), ("2", "1")] if "--quick" in sys.argv else [("2", "0"), ("2", "1"), ("1", "0"), ("0", "0")]
for trim, copies_only in MODES:  # what travels: rows the kernel reads/writes (default) / species only / whole fields
  os.environ["CLOUDSC_B200_E2E_TRIM"] = trim
  os.environ["CLOUDSC_B200_E2E_COPIES_ONLY"] = copies_only
  print(f"--- CLOUDSC_B200_E2E_TRIM={trim} COPIES_ONLY={copies_only}", flush=True)
  for chunk, ns in [(0, 0), (54, 4), (107, 3), (27, 4), (14, 8)]:
    ts = []
    for it in range(4):
        H["PLUDE"].copy_(plude0)
        t0 = time.perf_counter()
        st = api.run_host(H, ngptot, nproma, klev, pts, chunk, ns)
        ts.append(1e3 * (time.perf_counter() - t0))
    print("chunk_blocks %4d nstreams %d : %s ms (kernel sum %.1f ms, %.3f + %.3f GB)" % (chunk, ns, " ".join("%.1f" % x for x in ts), st["kernel_ms"], st["h2d_bytes"] / 1e9, st["d2h_bytes"] / 1e9), flush=True)
