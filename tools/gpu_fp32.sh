#!/bin/bash
# FP32 build: parity tests + timing (resident state, then caller-owned fields). usage: bash tools/gpu_fp32.sh
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "fp32" 2>&1 | tail -15
for np in 128; do
  timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --nproma $np --precision fp32 > gpurun_out/fp32_state_$np.json 2> gpurun_out/fp32_state_$np.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/fp32_state_$np.json"))
    print("fp32 nproma $np: ms/step %.3f  Mcol/s %.2f  frac %.3f  ok %s  %s" % (d["ms_per_step"], d["value"]/1e6, d["roofline"]["frac"], d["validated_vs_reference_h5"], d["kernel_info"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/fp32_state_$np.err").read()[-3000:])
PY
done
