#!/bin/bash
# time alternative builds of the library: bash tools/gpu_alt.sh <tag> [<tag> ...]   ("default" = the in-tree lib)
mkdir -p gpurun_out
for t in "$@"; do
  if [ "$t" = default ]; then unset CLOUDSC_B200_LIB; else export CLOUDSC_B200_LIB=$PWD/dwarf_p_cloudsc_b200/lib_alt/$t/libcloudsc_b200.so; fi
  timeout -s KILL 300 python bench.py --steps ${ALT_STEPS:-30} --warmup 3 --no-cpu --no-e2e --no-gpu-reference ${ALT_ARGS} > gpurun_out/alt_$t.json 2> gpurun_out/alt_$t.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/alt_$t.json")); print("alt $t: ms/step %.3f  Mcol/s %.2f  frac %.3f ok %s %s" % (d["ms_per_step"], d["value"]/1e6, d["roofline"]["frac"], d["validated_vs_reference_h5"], d["kernel_info"]))
except Exception as e:
    print("alt $t failed", e); print(open("gpurun_out/alt_$t.err").read()[-1500:])
PY
done
