#!/bin/bash
# parity + timing of every kernel variant. usage: bash tools/gpu_variants.sh "1 2 3" [pytest -k expr]
set -x
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for v in ${1:-1 2 3}; do
  timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --variant $v > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_v$v.json"))
    print("variant $v: ms/step %.3f  Mcol/s %.2f  frac %.3f  ok %s  regs %s" % (d["ms_per_step"], d["value"]/1e6, d["roofline"]["frac"], d["validated_vs_reference_h5"], d["kernel_info"]))
except Exception as e:
    print("variant $v failed", e); print(open("gpurun_out/bench_v$v.err").read()[-2000:])
PY
done
