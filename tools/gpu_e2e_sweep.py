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
# A/B on the SAME box (PCIe rates differ from box to box and drift with the neighbours' traffic): alternate the variants.
# (pipe, trim, copies_only): pipe = stream layout of run_host, trim = what travels, copies_only = kernels skipped
quick = "--quick" in sys.argv
VARIANTS = [("ring", "2", "0"), ("lanes", "2", "0"), ("ring", "2", "1"), ("lanes", "2", "1")]
if not quick:
    VARIANTS += [("ring", "1", "0"), ("ring", "0", "0")]
GEOMS = [(0, 0), (54, 4), (54, 6), (107, 3), (107, 4), (27, 8)]
best = {}
for rnd in range(2):
  for pipe, trim, copies_only in VARIANTS:
    os.environ["CLOUDSC_B200_E2E_PIPE"] = pipe
    os.environ["CLOUDSC_B200_E2E_TRIM"] = trim
    os.environ["CLOUDSC_B200_E2E_COPIES_ONLY"] = copies_only
    print(f"--- round {rnd}: pipe={pipe} trim={trim} copies_only={copies_only}", flush=True)
    for chunk, ns in GEOMS:
        ts = []
        for it in range(4):
            H["PLUDE"].copy_(plude0)
            t0 = time.perf_counter()
            st = api.run_host(H, ngptot, nproma, klev, pts, chunk, ns)
            ts.append(1e3 * (time.perf_counter() - t0))
        key = (pipe, trim, copies_only, chunk, ns)
        best[key] = min(best.get(key, 1e9), min(ts))
        print("chunk_blocks %4d lanes %d : %s ms (kernel sum %.1f ms, %.3f + %.3f GB)" % (chunk, ns, " ".join("%.1f" % x for x in ts), st["kernel_ms"], st["h2d_bytes"] / 1e9, st["d2h_bytes"] / 1e9), flush=True)
print("--- best of all repetitions [ms]")
for key in sorted(best, key=lambda k: best[k]):
    print("  pipe=%-5s trim=%s copies_only=%s chunk %4d lanes %d : %.1f" % (*key, best[key]))
