#!/bin/bash
# sweep of the phase stagger (ns between resident-warp slots): bash tools/gpu_stagger.sh "0 10000 20000 ..."
mkdir -p gpurun_out
for s in $1; do
  CLOUDSC_B200_STAGGER_NS=$s timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e ${2:-} > gpurun_out/stagger_$s.json 2> gpurun_out/stagger_$s.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/stagger_$s.json").read().strip().splitlines()[-1]); print("stagger $s: ms/step %.3f  Mcol/s %.2f  frac %.3f ok %s" % (d["ms_per_step"], d["value"]/1e6, d["roofline"]["frac"], d["validated_vs_reference_h5"]))
except Exception as e:
    print("stagger $s failed", e); print(open("gpurun_out/stagger_$s.err").read()[-1500:])
PY
done
