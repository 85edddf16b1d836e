#!/usr/bin/env python
"""Summarise one `ncu --set full` report (.ncu-rep) of the sweep kernel into a small text file for profiles/.
usage: ncu_summary.py <report.ncu-rep> <out.txt> [traffic.json]"""
import csv, io, json, re, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
pat = re.compile(
    r"^(Kernel Name|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic|occupancy_limit_\w+|waves_per_multiprocessor)|"
    r"sm__warps_active\.avg\.per_cycle_active|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|sm__inst_executed_pipe_fp64\.avg\.pct_of_peak_sustained_active|"
    r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|sass__inst_executed_local_(loads|stores)|"
    r"smsp__sass_inst_executed_op_shared_(ld|st)\.sum|l1tex__t_sector_hit_rate\.pct|lts__t_sector_hit_rate\.pct|"
    r"smsp__thread_inst_executed_per_inst_executed\.ratio|sm__cycles_elapsed\.avg\.per_second)$")
lines = []
traffic = None
for r in rows[2:]:
    d = dict(zip(hdr, r))
    lines.append("== launch %s: %s" % (d.get("ID", "?"), d.get("Kernel Name", "?")))
    rd = wr = None
    for h, u in zip(hdr, units):
        if pat.match(h) and h != "Kernel Name":
            lines.append("%-80s %-12s %s" % (h, u, d[h]))
            if h == "dram__bytes_read.sum": rd = (float(d[h]), u)
            if h == "dram__bytes_write.sum": wr = (float(d[h]), u)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    if rd and wr:
        traffic = rd[0] * scale[rd[1]] + wr[0] * scale[wr[1]]
        lines.append("dram traffic per launch (read+write): %.4f GB" % (traffic / 1e9))
open(out, "w").write("\n".join(lines) + "\n")
if len(sys.argv) > 3 and traffic:
    # keep what other captures recorded (per kernel); the top-level figure is the latest capture's
    try:
        rec = json.load(open(sys.argv[3]))
    except Exception:
        rec = {}
    m = re.search(r"cloudsc_sweep_\w+", d.get("Kernel Name", ""))
    kname = m.group(0) if m else d.get("Kernel Name", "?")
    by = rec.get("by_kernel", {})
    by[kname] = {"dram_bytes_per_launch": traffic, "source": rep.split("/")[-1]}
    rec.update({"dram_bytes_per_launch": traffic, "source": rep.split("/")[-1], "kernel": kname, "by_kernel": by})
    json.dump(rec, open(sys.argv[3], "w"), indent=1)
print("\n".join(lines))
