#!/usr/bin/env python
"""Per call-site line of cloudsc_column(): executed warp-instructions per warp-level and stall samples.

usage: ncu_by_section.py <sass.csv from `ncu --page source --csv --print-source sass`> <cubin built with -lineinfo> <kernel substring> [warp_levels]
Joins ncu's per-SASS-instruction counters (address order) with `nvdisasm -gi` inline chains and attributes every
instruction to the OUTERMOST line inside cloudsc_column.cuh (the statement of the column code it was inlined into),
so that mx()/exp_fast()/rcp_fast() are charged to the physics block that calls them.
"""
import csv, re, subprocess, sys, collections

def chains(cubin, kern):
    txt = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    out, active, chain = [], False, []
    fresh = False
    for ln in txt.splitlines():
        if ln.startswith("//-----") and ".text." in ln:
            active = kern in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            if not fresh:
                chain = []
                fresh = True
            chain.append((m.group(1).split("/")[-1], int(m.group(2))))
            if m.group(3):
                chain.append((m.group(3).split("/")[-1], int(m.group(4))))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            fresh = False
            out.append((int(m.group(1), 16), m.group(2).strip(), list(chain)))
    return out

def main():
    sass_csv, cubin, kern = sys.argv[1:4]
    warp_levels = float(sys.argv[4]) if len(sys.argv) > 4 else 5120 * 137
    rows = list(csv.reader(open(sass_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]; col = {n: i for i, n in enumerate(hdr)}
    data = rows[hi + 1:]
    ch = chains(cubin, kern)
    if len(ch) != len(data):
        print(f"warning: {len(ch)} disassembled vs {len(data)} profiled", file=sys.stderr)
    agg = collections.defaultdict(lambda: [0, 0, 0, 0, collections.Counter()])
    tot_e = tot_s = 0
    for (addr, text, chain), r in zip(ch, data):
        e = int(r[col["Instructions Executed"]] or 0); s = int(r[col["# Samples"]] or 0)
        key = None
        for f, l in reversed(chain):          # outermost first
            if f == "cloudsc_column.cuh":
                key = l; break
        if key is None:
            key = -(chain[-1][1] if chain else 0)  # kernel file lines as negatives
        a = agg[key]; a[0] += e; a[1] += s; a[2] += 1
        op = re.sub(r"^@!?U?P\d+\s+", "", text).split()[0]
        if op.startswith(("DFMA", "DADD", "DMUL", "DSETP")): a[3] += e
        for n in hdr:
            if n.startswith("stall_") and "Not Issued" not in n:
                v = int(r[col[n]] or 0)
                if v: a[4][n[6:]] += v
        tot_e += e; tot_s += s
    print(f"total warp-instructions {tot_e:,}  samples {tot_s:,}  per warp-level {tot_e / warp_levels:.0f}")
    print("line   inst/warp-level  inst%  fp64/wl  samp%  #sass  top stalls")
    for k in sorted(agg):
        e, s, n, f, st = agg[k]
        if e / tot_e < 0.001 and s / tot_s < 0.001: continue
        print(f"{k:6d} {e / warp_levels:10.1f} {100 * e / tot_e:7.2f} {f / warp_levels:8.1f} {100 * s / tot_s:6.2f} {n:6d}  " +
              ", ".join(f"{a} {b}" for a, b in st.most_common(3)))

if __name__ == "__main__":
    main()
