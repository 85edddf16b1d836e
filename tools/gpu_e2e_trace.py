"""Time line of cloudsc_b200_run_host calls (CLOUDSC_B200_E2E_TRACE=1) for the stream layouts and for three kinds of
page-locked host arrays.  usage: python tools/gpu_e2e_trace.py [--full]   (--full prints every chunk)"""
import os, re, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if "--child" not in sys.argv:
    for alloc in ("torch", "hostalloc", "thp"):
        for pipe in ("ring", "lanes"):
            env = dict(os.environ, CLOUDSC_B200_E2E_PIPE=pipe)
            r = subprocess.run([sys.executable, __file__, "--child", alloc], env=env, capture_output=True, text=True)
            rows = [l for l in r.stderr.splitlines() if re.match(r"\s+\d+\s+\d+ \|", l)]
            up = [float(x) for l in rows for x in re.findall(r"\|\s+([\d.]+) \.\.\s+([\d.]+) \|", l)[0]]
            dn = [tuple(float(x) for x in re.findall(r"\|\s+([\d.]+) \.\.\s+([\d.]+)\s*$", l)[0]) for l in rows]
            ups = [up[i + 1] - up[i] for i in range(0, len(up), 2)]
            dns = [b - a for a, b in dn]
            mid = slice(4, -2)
            print(f"alloc={alloc:9s} pipe={pipe:5s}: {[l for l in r.stderr.splitlines() if 'python wall' in l][-1].strip()};"
                  f" per full chunk: upload {sum(ups[mid]) / len(ups[mid]):.2f} ms, download {sum(dns[mid]) / len(dns[mid]):.2f} ms"
                  f" (first upload, alone: {ups[0]:.2f} ms)", flush=True)
            if "--full" in sys.argv:
                print(r.stderr)
    sys.exit(0)

import numpy as np
import torch
from dwarf_p_cloudsc_b200 import api, params, state
alloc = sys.argv[sys.argv.index("--child") + 1]
cols, scalars, _, klev = state.load_input()
p, k = params.params_from_scalars(scalars)
api.set_params(p, k)
ngptot, nproma = 163840, 128
pts = float(scalars["PTSPHY"])
F = state.blocked_fields(cols, ngptot, nproma, klev)
H, keep = {}, []
for n, a in F.items():
    if alloc == "torch":
        H[n] = torch.from_numpy(a).pin_memory()
    else:
        hb = api.HostBuffer(a.shape, a.dtype, api.HOST_HUGEPAGES if alloc == "thp" else 0)
        hb.array[...] = a
        keep.append(hb)
        H[n] = torch.from_numpy(hb.array)
plude0 = H["PLUDE"].clone()
os.environ["CLOUDSC_B200_E2E_TRACE"] = "0"
for _ in range(3):
    H["PLUDE"].copy_(plude0); api.run_host(H, ngptot, nproma, klev, pts)
best = 1e9
for _ in range(5):
    H["PLUDE"].copy_(plude0)
    t0 = time.perf_counter(); api.run_host(H, ngptot, nproma, klev, pts); best = min(best, 1e3 * (time.perf_counter() - t0))
os.environ["CLOUDSC_B200_E2E_TRACE"] = "1"
H["PLUDE"].copy_(plude0)
t0 = time.perf_counter()
st = api.run_host(H, ngptot, nproma, klev, pts)
print(f"python wall {1e3 * (time.perf_counter() - t0):.2f} ms (best of 5 untraced: {best:.2f}), library wall {st['wall_ms']:.2f} ms", file=sys.stderr, flush=True)
