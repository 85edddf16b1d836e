"""Experiment: the sweep kernel reading/writing PINNED HOST memory directly (UVA zero-copy over PCIe) instead of the
chunked copy pipeline of cloudsc_b200_run_host.   python tools/gpu_zero_copy.py [ngptot]
Modes: in=host/out=host, in=device/out=host (pure PCIe write path), in=host/out=device (pure PCIe read path)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dwarf_p_cloudsc_b200 import api, params, state, validate

ngptot = int(sys.argv[1]) if len(sys.argv) > 1 else 163840
nproma = 128
cols, scalars, _, klev = state.load_input()
p, k = params.params_from_scalars(scalars)
pts = float(scalars["PTSPHY"])
api.set_params(p, k)
F = state.blocked_fields(cols, ngptot, nproma, klev)
H = {n: torch.from_numpy(a).pin_memory() for n, a in F.items()}
D = {n: t.cuda() for n, t in H.items()}
plude0 = H["PLUDE"].clone()
ref = state.load_reference()


def use(n):
    return api.lib.cloudsc_b200_field_use(params.FIELD_INDEX[n])


def run(tag, pick):
    T = {n: pick(n) for n in H}
    ms = []
    for _ in range(4):
        H["PLUDE"].copy_(plude0)
        D["PLUDE"].copy_(plude0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        api.launch(T, ngptot, nproma, klev, pts)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    nv = 100
    errs = [validate.field_error(n, state.unblock(T[n][:1].cpu().numpy(), nv)[..., :nv], ref[n][..., :nv], nv)
            for n in state.VALIDATE_ORDER]
    ok = all(e.passes("fp64") for e in errs)
    print(f"{tag:34s} ms {['%.1f' % m for m in ms]}  valid {ok}", flush=True)


print("field use codes:", {n: use(n) for n in list(H)[:4]}, flush=True)
run("in=host out=host (zero-copy)", lambda n: H[n])
run("in=device out=host (PCIe writes)", lambda n: H[n] if use(n) == b"W" else D[n])
run("in=host out=device (PCIe reads)", lambda n: D[n] if use(n) == b"W" else H[n])
for _ in range(3):
    H["PLUDE"].copy_(plude0)
    t0 = time.perf_counter()
    api.run_host(H, ngptot, nproma, klev, pts)
    t1 = time.perf_counter()
    print("run_host pipeline: %.1f ms" % ((t1 - t0) * 1e3), flush=True)


def hybrid(nchunks):
    """inputs read by the kernel straight from pinned host memory (TMA bulk copies over PCIe), outputs to device
    buffers, copied back chunk by chunk by the copy engine while the next chunk computes"""
    nb = H["PT"].shape[0]
    per = (nb + nchunks - 1) // nchunks
    sk, sc = torch.cuda.Stream(), torch.cuda.Stream()
    H["PLUDE"].copy_(plude0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for c in range(nchunks):
        b0, b1 = c * per, min(nb, (c + 1) * per)
        if b0 >= b1:
            break
        T = {n: (D[n][b0:b1] if use(n) == b"W" else H[n][b0:b1]) for n in H}
        ncol = min(ngptot, b1 * nproma) - b0 * nproma
        with torch.cuda.stream(sk):
            api.launch(T, ncol, nproma, klev, pts, stream=sk.cuda_stream)
            ev = torch.cuda.Event()
            ev.record(sk)
        with torch.cuda.stream(sc):
            sc.wait_event(ev)
            for n in H:
                if use(n) == b"W":
                    H[n][b0:b1].copy_(D[n][b0:b1], non_blocking=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    nv = 100
    errs = [validate.field_error(n, state.unblock(H[n][:1].numpy(), nv)[..., :nv], ref[n][..., :nv], nv)
            for n in state.VALIDATE_ORDER]
    return (t1 - t0) * 1e3, all(e.passes("fp64") for e in errs)


for nch in (8, 16, 32, 16):
    ms, ok = hybrid(nch)
    print(f"hybrid (kernel reads host, CE copies outputs) {nch:3d} chunks: {ms:.1f} ms valid {ok}", flush=True)
