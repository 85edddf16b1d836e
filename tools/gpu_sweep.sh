#!/bin/bash
# BASELINE configs[3] and [4]: NPROMA sweep 32..256 and the FP32 build at NGPTOT=163840, plus the
# NGPTOT=655360 end-to-end run.  Prints one summary line per run; JSON lines in gpurun_out/sweep_*.json
set -x
mkdir -p gpurun_out
for np in 32 64 128 256; do
  timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-gpu-reference --nproma $np > gpurun_out/sweep_fp64_nproma$np.json 2> gpurun_out/sweep_fp64_nproma$np.err
done
for np in 32 64 128 256; do
  timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-gpu-reference --nproma $np --precision fp32 > gpurun_out/sweep_fp32_nproma$np.json 2> gpurun_out/sweep_fp32_nproma$np.err
done
timeout -s KILL 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-gpu-reference --ngptot 655360 --e2e-steps 3 > gpurun_out/sweep_e2e_655360.json 2> gpurun_out/sweep_e2e_655360.err
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-gpu-reference --precision fp32 --e2e-steps 3 > gpurun_out/sweep_fp32_e2e.json 2> gpurun_out/sweep_fp32_e2e.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/sweep_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e")
        print("%-44s ms %.3f  Mcol/s %.2f  frac %.3f  valid %s  e2e %s" % (f, d["ms_per_step"], d["value"] / 1e6, d["roofline"]["frac"],
              d["validated_vs_reference_h5"], ("%.3f Mcol/s %.1f ms" % (e["value"] / 1e6, e["ms_per_step"])) if e else None))
    except Exception as ex:
        print(f, "failed", ex); print(open(f.replace(".json", ".err")).read()[-1200:])
PY
