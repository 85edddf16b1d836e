#!/bin/bash
# instruction-cache pressure of kernel variants / alternative builds:
#   bash tools/gpu_icache.sh "<variant ids>" [<alt tag> ...]      (alt tags: dwarf_p_cloudsc_b200/lib_alt/<tag>)
# per configuration: a plain timing run, then ONE ncu pass with the instruction-cache metrics of one sweep launch
mkdir -p gpurun_out
M=sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
run() {  # name, extra bench args
  local name=$1; shift
  timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e "$@" > gpurun_out/ic_$name.json 2> gpurun_out/ic_$name.err || { echo "$name: bench failed"; tail -5 gpurun_out/ic_$name.err; return; }
  timeout -s KILL 300 ncu --metrics $M --clock-control none -k regex:cloudsc_sweep -s 2 -c 1 --csv --log-file gpurun_out/ic_$name.csv \
      python bench.py --steps 1 --warmup 2 --no-cpu --no-e2e "$@" > gpurun_out/ic_${name}_ncu.log 2>&1
  python - <<PY
import json, csv
d = json.load(open("gpurun_out/ic_$name.json"))
m = {}
try:
    rows = [r for r in csv.reader(open("gpurun_out/ic_$name.csv")) if len(r) > 5]
    h = rows[0]
    for r in rows[1:]:
        m[r[h.index("Metric Name")]] = r[h.index("Metric Value")]
except Exception as e:
    m = {"err": str(e)}
print("%-14s ms %.3f frac %.3f ok %s regs %d | icc hit %s%% gcc-inst %s (%s%% of peak) inst %s no_inst/issue %s issue %s%%" % (
    "$name", d["ms_per_step"], d["roofline"]["frac"], d["validated_vs_reference_h5"], d["kernel_info"]["registers"],
    m.get("sm__icc_request_hit_rate.pct"), m.get("gcc__cache_requests_type_instruction.sum"),
    m.get("gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed"), m.get("smsp__inst_executed.sum"),
    m.get("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"), m.get("smsp__issue_active.avg.pct_of_peak_sustained_active")))
PY
}
for v in $1; do run v$v --variant $v; done
shift
for t in "$@"; do
  export CLOUDSC_B200_LIB=$PWD/dwarf_p_cloudsc_b200/lib_alt/$t/libcloudsc_b200.so
  run alt_$t
  unset CLOUDSC_B200_LIB
done
