#!/bin/bash
# direct kernel (variant 2) with L1 prefetch distance d: bash tools/gpu_prefetch.sh "0 1 2 3"
mkdir -p gpurun_out
for s in $1; do
  CLOUDSC_B200_PREFETCH=$s timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --variant 2 > gpurun_out/prefetch_$s.json 2> gpurun_out/prefetch_$s.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/prefetch_$s.json").read().strip().splitlines()[-1]); print("prefetch $s: ms/step %.3f  Mcol/s %.2f  frac %.3f ok %s" % (d["ms_per_step"], d["value"]/1e6, d["roofline"]["frac"], d["validated_vs_reference_h5"]))
except Exception as e:
    print("prefetch $s failed", e); print(open("gpurun_out/prefetch_$s.err").read()[-1500:])
PY
done
