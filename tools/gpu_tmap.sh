#!/bin/bash
# first contact of the tensor-map kernel: its parity tests (short timeout: a wrong TMA descriptor can hang), then bench
set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -m gpu -x -q -k "arena_state or state_handle" > gpurun_out/pytest_tmap.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tmap.log
tail -30 gpurun_out/pytest_tmap.log
grep -q "rc=0" gpurun_out/pytest_tmap.log || exit 1
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_tmap.json 2> gpurun_out/bench_tmap.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_tmap.json").read().strip().splitlines()[-1]); print("tmap: ms/step %.3f  Mcol/s %.2f  frac %.3f ok %s %s %s" % (d["ms_per_step"], d["value"]/1e6, d["roofline"]["frac"], d["validated_vs_reference_h5"], d["kernel_info"], d["device_api"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_tmap.err").read()[-2500:])
PY
