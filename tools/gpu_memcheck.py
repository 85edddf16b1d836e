"""Tiny driver for `compute-sanitizer --tool memcheck`: every kernel family once on ragged geometries."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dwarf_p_cloudsc_b200 import api, params, state

cols, scalars, _, klev = state.load_input()
p, k = params.params_from_scalars(scalars)
api.set_params(p, k)
for variant, ngptot, nproma in [(0, 257, 32), (0, 300, 128), (5, 384, 128), (8, 200, 64), (9, 77, 16), (2, 100, 100), (0, 64, 64)]:
    api.set_variant(variant)
    F = state.blocked_fields(cols, ngptot, nproma, klev)
    D = {n: torch.from_numpy(a).cuda() for n, a in F.items()}
    api.launch(D, ngptot, nproma, klev, float(scalars["PTSPHY"]))
    torch.cuda.synchronize()
    print("variant", variant, ngptot, nproma, "ok", float(D["PFPLSN"].abs().sum()))
H = {n: torch.from_numpy(a).pin_memory() for n, a in state.blocked_fields(cols, 1000, 64, klev).items()}
api.set_variant(0)
print(api.run_host(H, 1000, 64, klev, float(scalars["PTSPHY"]), chunk_blocks=3, nstreams=2))
