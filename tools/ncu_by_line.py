#!/usr/bin/env python
"""Aggregate an ncu `--page source --csv --print-source sass` dump by CUDA source line.

usage: ncu_by_line.py <sass.csv> <object-or-cubin> <kernel-substring> [top]
Joins ncu's per-SASS-instruction counters (in address order) with `nvdisasm -g` line info of the same
kernel, then prints executed warp-instructions and stall samples per source line / per inlined function.
"""
import csv, re, subprocess, sys, tempfile, os, collections

def line_map(obj, kern):
    tmp = tempfile.mkdtemp()
    if obj.endswith(".cubin"):
        cub = obj
    else:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
        cub = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    out, cur, active = [], ("?", 0), False
    for ln in txt.splitlines():
        if ln.startswith("//-----") and ".text." in ln:
            active = kern in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return out

def main():
    sass_csv, obj, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(sass_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    data = rows[hdr_i + 1:]
    lm = line_map(obj, kern)
    if len(lm) != len(data):
        print(f"warning: {len(lm)} disassembled vs {len(data)} profiled instructions", file=sys.stderr)
    agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    tot_inst = tot_samp = 0
    opc = collections.Counter()
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    stalls = collections.Counter()
    for (addr, text, src), r in zip(lm, data):
        inst = int(r[col["Instructions Executed"]] or 0)
        samp = int(r[col["# Samples"]] or 0)
        a = agg[src]
        a[0] += inst; a[1] += samp; a[2] += 1
        tot_inst += inst; tot_samp += samp
        op = re.sub(r"^@!?U?P\d+\s+", "", text).split()[0].split(".")[0]
        opc[op] += inst
        for s in stall_cols:
            v = int(r[col[s]] or 0)
            stalls[s] += v
            a[3][s] += v
    print(f"total warp-instructions {tot_inst:,}  samples {tot_samp:,}")
    print("opcode mix (executed):", ", ".join(f"{k} {100*v/tot_inst:.1f}%" for k, v in opc.most_common(18)))
    print("stall mix:", ", ".join(f"{k[6:]} {100*v/max(1,sum(stalls.values())):.1f}%" for k, v in stalls.most_common(8)))
    print(f"{'file:line':38s} {'inst%':>7s} {'samp%':>7s} {'#sass':>6s}  top stalls")
    for src, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        ts = ", ".join(f"{k[6:]} {v}" for k, v in a[3].most_common(3))
        print(f"{src[0]+':'+str(src[1]):38s} {100*a[0]/tot_inst:7.2f} {100*a[1]/max(1,tot_samp):7.2f} {a[2]:6d}  {ts}")

main()
