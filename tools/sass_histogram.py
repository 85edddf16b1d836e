"""SASS opcode histogram per kernel of libcloudsc_b200.so (cuobjdump -sass), so that the Blackwell-specific instructions
(UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk, SYNCS = mbarrier, ELECT, LDGSTS = cp.async, STG.E.EF = streaming
stores, MUFU.RCP64H / RSQ64H seeds) are on record.  usage: python tools/sass_histogram.py [lib] > profiles/rN_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dwarf_p_cloudsc_b200", "lib", "libcloudsc_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\n\s*(REG:\d+ STACK:\d+ SHARED:\d+ LOCAL:\d+)", res):
    usage[m.group(1)] = m.group(2)
kern, hist = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1); hist[kern] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ("UTMALDG", "UBLKCP", "SYNCS", "ELECT", "LDGSTS", "STG.E.EF", "MUFU", "LDS", "STS", "LDG", "DFMA", "DADD", "DMUL", "DSETP", "FSEL", "BRA", "CALL")
print(f"# {os.path.relpath(lib, ROOT)}: SASS opcode histogram per kernel (cuobjdump -sass, static counts)")
for kern in sorted(hist):
    h = hist[kern]
    name = subprocess.run(["c++filt", kern], capture_output=True, text=True).stdout.strip().split("(")[0]
    total = sum(h.values())
    print(f"\n== {name}   [{usage.get(kern, '')}]  {total} instructions, {total * 16 / 1024:.1f} KB")
    fam = collections.Counter()
    for op, n in h.items():
        for k in KEY:
            if op.startswith(k):
                fam[k] += n
    print("   key families: " + ", ".join(f"{k} {fam[k]}" for k in KEY if fam[k]))
    print("   " + ", ".join(f"{op} {n}" for op, n in h.most_common(28)))
    special = [f"{op} {n}" for op, n in sorted(h.items()) if any(op.startswith(k) for k in ("UTMALDG", "UBLKCP", "SYNCS", "ELECT", "LDGSTS", "MUFU", "STG"))]
    print("   Blackwell / memory-path opcodes: " + ", ".join(special))
