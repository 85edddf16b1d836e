#!/usr/bin/env python
"""bench.py -- CLOUDSC columns/s (and GFLOP/s, dwarf flop model) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): FP64, NGPTOT = 163 840 columns per GPU, NPROMA = 128, KLEV = 137,
the 100 reference columns replicated (g -> g mod 100), fields resident in HBM.  A "step" is one pass of
the hot path (the whole block loop of cloudsc_driver.c:183-229 = ONE kernel launch) over that batch.
PLUDE is in/out (cloudsc_c.c:970), so it is restored from a pristine copy between steps, outside the
per-step CUDA-event pairs.  Multi-GPU: every rank owns a contiguous range of columns, no collective on
the data path; time = max over ranks, value = sum of columns / that time ("weak" scaling).

The one JSON line printed by rank 0 carries value / e2e / roofline / cpu_baseline / clocks / gpu_launches
as the contract in the task statement describes.  `--impl reference` times the reference's own CPU
implementation (oracle/_ref = UNMODIFIED cloudsc_c.c; else the C port) on the host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NGPTOT = 163840
NPROMA = 128
METRIC = "cloudsc_columns_per_s_fp64"
UNIT = "columns/s"
CPU_NPROMA = 32   # the CPU-friendly blocking of configs[0] (benchmark/include/include_run.yml:35-36)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on
# stdout at the first collective), so file descriptor 1 is pointed at stderr for the whole run and the
# result line goes to a private duplicate of the original stdout.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region: the sampler runs from before
    the warm-up (nvidia-smi needs ~0.1 s to deliver its first row) and only rows time-stamped inside
    [mark_begin, mark_end] are summarised."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        import datetime
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = self.t0 or 0.0
        t1 = self.t1 or time.time()

        def stamp(row_time, text):
            try:   # nvidia-smi's own clock (same host), e.g. 2026/10/18 12:34:56.789
                return datetime.datetime.strptime(text, "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except Exception:
                return row_time
        parsed = []
        for rt, r in self.rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 8:
                continue
            parsed.append((stamp(rt, c[0]), c[1:]))
        inside = [c for ts, c in parsed if t0 <= ts <= t1 + 0.02]
        if not inside:                       # very short region: the rows closest after its start
            inside = [c for ts, c in parsed if ts >= t0][:2] or [c for _, c in parsed][-2:]
        for c in inside:
            try:
                sm.append(float(c[0])); mx.append(float(c[1])); pw.append(float(c[2]))
            except ValueError:
                continue
            for n, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def workload_name(ngptot, nproma, klev, precision):
    return f"CLOUDSC {precision.upper()} NGPTOT={ngptot} per GPU, NPROMA={nproma}, KLEV={klev}, 100 reference columns replicated"


def cpu_arm(ngptot_target, seconds_target=12.0, nthreads=0):
    """The CPU checkers to time on a bounded sample: the parity build of the unmodified reference kernel (-O2,
    contraction off; else the port) and, where the host can run it, the same source built for speed (-O3
    -march=x86-64-v4/-v3, oracle/Makefile `reffast`)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    from dwarf_p_cloudsc_b200 import params, state
    os.environ.setdefault("OMP_SCHEDULE", "static")
    os.environ.setdefault("OMP_PROC_BIND", "close")
    try:
        orc = pyoracle.best_oracle()
    except Exception:
        pyoracle.build(("port",))
        orc = pyoracle.best_oracle()
    builds = [orc]
    try:
        builds.append(pyoracle.RefFastOracle())
    except Exception as exc:
        log(f"no fast reference build on this host: {exc}")
    cores = nthreads or (os.cpu_count() or 1)
    cols, scalars, _, klev = state.load_input()
    p, k = params.params_from_scalars(scalars)
    pts = scalars["PTSPHY"]

    def once(n, which=orc):
        F = state.blocked_fields(cols, n, CPU_NPROMA, klev)
        return which.run(F, p, k, n, CPU_NPROMA, klev, pts, cores)

    n0 = min(ngptot_target, 64 * cores)
    once(n0)                    # warm-up (page faults, thread pool)
    t0 = once(n0)
    rate = n0 / max(t0, 1e-9)
    n = int(min(ngptot_target, max(n0, rate * seconds_target)))
    n -= n % CPU_NPROMA
    n = max(n, CPU_NPROMA)
    return builds, klev, cores, n, once


def cpu_builds_line(builds, times, n):
    """{flags: columns/s} of every CPU build timed, and the fastest one."""
    rates = {getattr(b, "flags", b.kind): n / t for b, t in zip(builds, times)}
    best = max(range(len(builds)), key=lambda i: n / times[i])
    return rates, builds[best], times[best]


def run_reference(args, rank):
    if rank != 0:
        return
    # bounded sample: a step is at most the full 163 840 columns (~0.7 s on 16 cores) and the whole run ~45 s per build
    builds, klev, cores, n, once = cpu_arm(NGPTOT, seconds_target=min(4.0, max(0.05, 45.0 / max(args.steps, 1))))
    per_build = []
    for b in builds:
        for _ in range(max(args.warmup, 1)):
            once(n, b)
        ts = [once(n, b) for _ in range(args.steps)]
        per_build.append(sum(ts) / len(ts))
    rates, best, t = cpu_builds_line(builds, per_build, n)
    value = n / t
    from dwarf_p_cloudsc_b200.params import FLOP_PER_COLUMN
    sample = (f"{n} of {NGPTOT} columns per step (NPROMA={CPU_NPROMA}, {cores} OpenMP threads, schedule static), FP64, 100 reference "
              f"columns replicated; unmodified cloudsc_c.c built {getattr(best, 'flags', best.kind)} (the fastest of the builds timed)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (100 reference columns replicated)",
        "config": {"workload": workload_name(NGPTOT, NPROMA, klev, "fp64"), "ngptot_per_gpu": NGPTOT, "nproma": NPROMA, "klev": klev,
                   "reference_arm": f"dwarf-cloudsc-c kernel (unmodified cloudsc_c.c) on the host CPU, bounded sample of {n} columns per step, NPROMA={CPU_NPROMA}",
                   "sample_columns": n, "cpu_nproma": CPU_NPROMA},
        "gflops": 1e-9 * FLOP_PER_COLUMN * value,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": best.kind, "sample": sample, "builds_columns_per_s": rates},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def time_gpu_reference(args, H, plude_host0, p, k, ngptot, nproma, pts, st, our_ms):
    """The reference's own CUDA kernels (oracle/_ref/libcloudsc_gpuref_*.so = src/cloudsc_cuda/cloudsc/cloudsc_c_k_caching.cu
    and cloudsc_c.cu, unmodified, compiled for sm_100a), launched as cloudsc_driver.cu:391-416 launches them, on a device
    copy of the same fields, timed like the product kernel (CUDA events per launch, PLUDE restored outside them), their
    outputs compared with the product kernel's.  Measurement infrastructure only (oracle/), never the thing shipped."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from dwarf_p_cloudsc_b200 import state, validate
    out = {}
    try:
        import pyoracle
        ours = None
        for which in ("kcaching", "scc"):
            try:
                g = pyoracle.GpuRef(which)
            except FileNotFoundError:
                out[which] = {"unavailable": "oracle/_ref/%s not in this snapshot" % pyoracle.GpuRef.libs[which]}
                continue
            H["PLUDE"].copy_(plude_host0)
            D = {n: t.cuda() for n, t in H.items()}
            p0 = D["PLUDE"].clone()
            r = g.run(D, p, k, ngptot, nproma, pts, plude0=p0, warmup=2, reps=max(1, args.gpu_reference_steps))
            torch.cuda.synchronize()
            if ours is None:                                   # the product kernel's outputs of the same inputs
                ours = {n: H[n].numpy().copy() for n in state.VALIDATE_ORDER}
                st.reset_inout(); st.step(pts); st.download({n: torch.from_numpy(a) for n, a in ours.items()}); st.sync()
            errs = [validate.field_error(n, D[n].cpu().numpy(), ours[n], ngptot) for n in state.VALIDATE_ORDER]
            info = g.kernel_info()
            out[which] = {"kernel": g.source + " (unmodified, nvcc -O3 sm_100a)", "launch": f"block=({nproma},1,1) grid=(1,1,{(ngptot + nproma - 1) // nproma})",
                          "ms_per_step": r["ms"], "ms_min": r["ms_min"], "ms_max": r["ms_max"], "steps": max(1, args.gpu_reference_steps),
                          "columns_per_s": ngptot / (r["ms"] * 1e-3), "registers": info["registers"], "local_bytes_per_thread": info["local_bytes"],
                          "speedup_of_this_repo": r["ms"] / our_ms, "agrees_with_this_repo_fp64_bar": all(e.passes() for e in errs),
                          "worst_l1_vs_this_repo": max(e.rel_l1 for e in errs)}
            del D, p0
            torch.cuda.empty_cache()
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f).get("gpu_reference") or {}
            for which, d in tj.items():
                if which in out and "ms_per_step" in out[which]:
                    out[which]["dram_bytes"] = d.get("dram_bytes_per_launch")
        except Exception:
            pass
    except Exception as exc:   # a reported comparison, never allowed to break the bench line
        out["error"] = f"{type(exc).__name__}: {exc}"
    return out


def run_b200(args, rank, local_rank, world):
    import numpy as np
    import torch
    from dwarf_p_cloudsc_b200 import api, params, shard, state, validate

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the CLOUDSC hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if args.bind and world > 1:
        # one slice of the host cores per rank, set BEFORE the host arrays are allocated and first touched
        ncpu = os.cpu_count() or 1
        per = max(1, ncpu // world)
        try:
            os.sched_setaffinity(0, set(range(local_rank * per, min(ncpu, (local_rank + 1) * per))))
        except OSError as exc:
            log(f"[rank {rank}] sched_setaffinity failed: {exc}")
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)

    ngptot, nproma = args.ngptot, args.nproma
    fp32 = args.precision == "fp32"
    prec = api.FP32 if fp32 else api.FP64
    cols, scalars, _, klev = state.load_input()
    p, k = params.params_from_scalars(scalars)
    pts = float(scalars["PTSPHY"])
    api.set_params(p, k)
    if args.variant:
        api.set_variant(args.variant)

    # -- this rank's shard: global columns [rank*ngptot, (rank+1)*ngptot)  
    t_setup = time.time()
    first_column, _ = shard.weak_shard(rank, world, ngptot)
    F = state.blocked_fields(cols, ngptot, nproma, klev, first_column=first_column)
    H, keep = {}, []
    for n in list(F):                                                   # pinned host copies (e2e inputs), one field at a time
        a = F.pop(n)
        if args.pin == "torch":
            H[n] = torch.from_numpy(a).pin_memory()
        else:                   # the library's own allocator: plain cudaHostAlloc, or 2 MB pages first-touched by this rank
            hb = api.HostBuffer(a.shape, a.dtype, api.HOST_HUGEPAGES if args.pin == "thp" else 0)
            hb.array[...] = a
            keep.append(hb)
            H[n] = torch.from_numpy(hb.array)
    del F
    stream = torch.cuda.current_stream().cuda_stream
    # resident device state of the library (cloudsc_b200_create/upload/step): one arena, tensor-map TMA.
    # FP32 build (BASELINE configs[4]): the same API on float host arrays.
    if fp32:
        H = {n: (t.float().pin_memory() if t.dtype == torch.float64 else t) for n, t in H.items()}
    plude_host0 = H["PLUDE"].clone()
    st = api.State(ngptot, nproma, klev, prec)
    st.upload(H, stream)
    torch.cuda.synchronize()
    kinfo = st.kernel_info()
    sched = st.get_schedule()             # FP64 states default to the split-tail launch (mode 3) when the launch is big enough

    def restore():
        st.reset_inout(stream)          # PLUDE is in/out: pristine copy kept by upload()

    def step():
        st.step(pts, 0, 0, stream)
    log(f"[rank {rank}] setup {time.time() - t_setup:.1f}s; kernel {kinfo}; schedule {sched}")

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        restore()
        step()
    torch.cuda.synchronize()

    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = api.launch_counts()
    sampler.mark_begin()
    t_region = time.time()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        restore()                         # restore the in/out field, outside the event pair
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    launches1 = api.launch_counts()
    if dist:
        dist.barrier()
    # nvidia-smi needs ~0.1 s per row: keep the SAME step running (untimed) until the sampled window is >= 1.2 s long,
    # so that the clocks / throttle reasons reported are those of this kernel under sustained load
    sustain = 0
    while rank == 0 and time.time() - t_region < 1.2:
        restore()
        step()
        sustain += 1
        if sustain % 16 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = f"{args.steps} timed steps + {sustain} untimed steps of the same kernel"
    times_ms = [a.elapsed_time(b) for a, b in ev]
    ms = sum(times_ms) / len(times_ms)
    ms = shard.reduce_max(ms, dist, dev)          # time = max over ranks; no collective on the data path
    gpu_launches = sum(launches1.values()) - sum(launches0.values())
    launched = {n: launches1[n] - launches0[n] for n in launches1 if launches1[n] != launches0[n]}
    if "tmap" in launched:                # the counter is per kernel family; the schedule says which tensor-map kernel it was
        launched[KERNEL_OF_SCHEDULE[sched["used"]]] = launched.pop("tmap")

    # -- correctness of what was just timed: EVERY column of EVERY rank against the replicated reference.h5, the
    #    per-rank norms combined as the reference's MPI mode combines them (validate_mod.F90:197-199: MIN/MAX/SUM)
    ref = state.load_reference()
    restore()
    step()                                # (the sustain loop above may have left another step's PLUDE behind on rank 0)
    st.download(H, stream)                # outputs back in the pinned host arrays
    torch.cuda.synchronize()
    t_val = time.time()
    errs_local = validate.validate_replicated({n: H[n].numpy() for n in state.VALIDATE_ORDER}, ref, ngptot, nproma, first_column)
    errs = shard.reduce_field_errors(errs_local, world * ngptot, dist, dev)
    ok = all(e.passes(args.precision) for e in errs)
    validated_columns = world * ngptot
    log(f"[rank {rank}] validated {ngptot} columns in {time.time() - t_val:.1f}s: {'PASS' if all(e.passes(args.precision) for e in errs_local) else 'FAIL'}")
    if rank == 0 and not ok:
        log(validate.format_table(errs, args.precision))
    worst_l1 = max(e.rel_l1 for e in errs)

    # -- end to end through the public host API: pinned host buffers in, host buffers out
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        H["PLUDE"].copy_(plude_host0)
        for _ in range(2):                                    # warm-up (staging arenas, streams, first touch)
            H["PLUDE"].copy_(plude_host0)
            api.run_host(H, ngptot, nproma, klev, pts, precision=prec)
        e2e_ms, h2d, d2h, kern_ms = [], 0, 0, 0.0
        for _ in range(e2e_steps):
            H["PLUDE"].copy_(plude_host0)
            if dist:
                dist.barrier()                                # every rank starts the step together: a stationary sample
            t0 = time.perf_counter()
            rs = api.run_host(H, ngptot, nproma, klev, pts, precision=prec)
            e2e_ms.append(1e3 * (time.perf_counter() - t0))
            h2d, d2h, kern_ms = int(rs["h2d_bytes"]), int(rs["d2h_bytes"]), float(rs["kernel_ms"])
        # the e2e results must be the resident path's: every column against the replicated reference again
        e2e_errs = shard.reduce_field_errors(
            validate.validate_replicated({n: H[n].numpy() for n in state.VALIDATE_ORDER}, ref, ngptot, nproma, first_column),
            world * ngptot, dist, dev)
        e2e_ok = all(e.passes(args.precision) for e in e2e_errs)
        # per step: the slowest rank (every rank's sample of step i is the same barrier-to-finish interval)
        per_step = [shard.reduce_max(x, dist, dev) for x in e2e_ms]
        e2e_t = statistics.median(per_step)    # PCIe on a shared host: the median of the steps, all samples reported
        e2e = {"value": world * ngptot / (e2e_t * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "ms_per_step": e2e_t, "steps": e2e_steps,
               "step_ms": [round(x, 2) for x in per_step], "kernel_ms_sum_rank0": round(kern_ms, 3),
               "gb_per_s_per_gpu": (h2d + d2h) / (e2e_t * 1e-3) / 1e9, "validated_vs_reference_h5": e2e_ok,
               "host_memory": {"torch": "torch pin_memory (cudaHostAlloc)", "hostalloc": "cloudsc_b200_host_alloc_ex(0)",
                               "thp": "cloudsc_b200_host_alloc_ex(HUGEPAGES): 2 MB pages, first touch by the rank, cudaHostRegister"}[args.pin],
               "rank_core_binding": bool(args.bind and world > 1),
               "api": "cloudsc_b200_run_host (C ABI, pinned host fields, chunked multi-stream H2D|kernel|D2H; only the rows the kernel reads/writes travel)"}

    # -- the reference's own CUDA kernels on the same device, same fields, same timing protocol (N=1 only)
    gpu_reference = None
    if world == 1 and not args.no_gpu_reference and not fp32 and klev == 137:
        gpu_reference = time_gpu_reference(args, H, plude_host0, p, k, ngptot, nproma, pts, st, ms)

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    bytes_col = params.algorithmic_bytes_per_column(klev, 4 if fp32 else 8)
    peak, peak_src = measured_peak_gbs()
    value = world * ngptot / (ms * 1e-3)
    achieved = bytes_col * ngptot / (ms * 1e-3) / 1e9         # per GPU, the dominant (only) kernel
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)
            kname = KERNEL_OF_SCHEDULE[sched["used"]]
            if fp32:
                traffic = (traffic.get("fp32") or {}).get("dram_bytes_per_launch")
            elif kname in traffic.get("by_kernel", {}):
                traffic = traffic["by_kernel"][kname]["dram_bytes_per_launch"]
            else:                         # a capture of another kernel says nothing about this one
                traffic = traffic.get("dram_bytes_per_launch") if traffic.get("kernel", kname) == kname else None
        if ngptot != 163840:
            traffic = None                # the ncu capture is of the default workload
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": KERNEL_OF_SCHEDULE[sched["used"]],
                "algorithmic_bytes_per_column": bytes_col, "columns_per_launch": ngptot}

    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            builds, _kl, cores, n, once = cpu_arm(NGPTOT, seconds_target=8.0)
            times = [once(n, b) for b in builds]
            rates, best, t = cpu_builds_line(builds, times, n)
            cpu = {"value": n / t, "unit": UNIT, "cores": cores, "kind": best.kind, "builds_columns_per_s": rates,
                   "sample": f"{n} of {NGPTOT} columns, NPROMA={CPU_NPROMA}, {cores} OpenMP threads, one pass per build ({t:.2f} s); "
                             f"unmodified cloudsc_c.c built {getattr(best, 'flags', best.kind)} (the fastest build timed)"}
        except Exception as exc:  # the baseline is reported, never allowed to break the bench line
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {exc}"}

    line = {
        "metric": METRIC if not fp32 else METRIC.replace("fp64", "fp32"), "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if fp32 else "f64",
        "data": "synthetic (100 reference columns replicated to NGPTOT, g -> g mod 100)",
        "config": {"workload": workload_name(ngptot, nproma, klev, args.precision), "residency": "device-resident fields, one launch per step",
                   "ngptot_per_gpu": ngptot, "nproma": nproma, "klev": klev, "parallelism": f"columns sharded x{world}, no collective",
                   "l2": f"inputs_exceed_l2 ({bytes_col * ngptot / 1e9:.1f} GB of fields per step vs 126 MB L2)"},
        "gflops": 1e-9 * params.FLOP_PER_COLUMN * value,
        "e2e": e2e, "gpu_launches": gpu_launches, "gpu_launches_by_kernel": launched, "clocks": clocks, "roofline": roofline,
        "cpu_baseline": cpu, "gpu_reference": gpu_reference,
        "validated_vs_reference_h5": ok, "validated_columns": validated_columns, "validation_worst_l1": worst_l1,
        "device_api": "cloudsc_b200_create/upload/step (arena state)",
        "step_ms": {"min": min(times_ms), "median": statistics.median(times_ms), "max": max(times_ms)}, "kernel_info": kinfo, "schedule": sched,
    }
    emit(line)
    if dist:
        dist.destroy_process_group()


KERNEL_OF_SCHEDULE = {0: "cloudsc_sweep_tmap", 1: "cloudsc_sweep_persist", 2: "cloudsc_sweep_stream", 3: "cloudsc_sweep_split"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ngptot", type=int, default=NGPTOT, help="columns per GPU")
    ap.add_argument("--nproma", type=int, default=NPROMA)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"], help="fp32 = the single-precision build (BASELINE configs[4])")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (kernel sweeps only)")
    ap.add_argument("--pin", default="torch", choices=["torch", "hostalloc", "thp"],
                    help="page-locked host arrays of the e2e leg: torch pin_memory, the library's cudaHostAlloc, or 2 MB pages + cudaHostRegister")
    ap.add_argument("--bind", type=int, default=0, help="1: pin every rank to its own slice of the host cores before allocating")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip timing the reference's own CUDA kernels (N=1 leg)")
    ap.add_argument("--gpu-reference-steps", type=int, default=10)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not (world == 1 and args.gpus > 1 and args.impl == "b200"):   # (the re-launching parent must pass stdout through)
        capture_stdout()
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun when called as plain `python bench.py --gpus N`
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29541"),
               os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
