"""One launch of each of the reference's own CUDA kernels (oracle/_ref/libcloudsc_gpuref_*.so) at the bench size, for an
ncu pass:  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:cloudsc_c ... python tools/gpu_ref_run.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import pyoracle
from dwarf_p_cloudsc_b200 import params, state
ngptot, nproma = 163840, 128
cols, scalars, _, klev = state.load_input()
p, k = params.params_from_scalars(scalars)
F = state.blocked_fields(cols, ngptot, nproma, klev)
for which in ("kcaching", "scc"):
    D = {n: torch.from_numpy(a).cuda() for n, a in F.items()}
    r = pyoracle.GpuRef(which).run(D, p, k, ngptot, nproma, float(scalars["PTSPHY"]))
    torch.cuda.synchronize()
    print(which, r, flush=True)
    del D
