# per-device-function totals of an `ncu --page source --csv` export (address ranges from the ELF symbol table)
#   python tools/dev/ncu_func.py <source.csv> <kernel-name-fragment> [object]
import csv, sys, re, subprocess, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if r[ix["# Samples"]].isdigit()]
base = int(data[0][ix["Address"]], 16)
obj = sys.argv[3] if len(sys.argv) > 3 else "trajectory_planner_b200/csrc/build/tp_vigo.o"
out = subprocess.run(["cuobjdump", "-elf", obj], capture_output=True, text=True).stdout
syms = []
for l in out.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+|0)\s+0x2\s+\S+\s+\S+\s+\$(\S+?)\$(\S+)", l)
    if m and sys.argv[2] in m.group(3):
        syms.append((int(m.group(1), 16), int(m.group(2), 16), m.group(4)))
syms.sort()
def fn(off):
    for o, s, n in syms:
        if o <= off < o + s: return n
    return "(kernel body)"
samp = collections.Counter(); execd = collections.Counter(); ninst = collections.Counter()
for r in data:
    f = fn(int(r[ix["Address"]], 16) - base)
    samp[f] += int(r[ix["# Samples"]]); execd[f] += int(r[ix["Instructions Executed"]]); ninst[f] += 1
tot = sum(samp.values()); te = sum(execd.values())
print("%-60s %8s %6s %12s %6s %6s" % ("function", "samples", "%", "warp-instr", "%", "static"))
for f, s in samp.most_common():
    print("%-60s %8d %6.2f %12d %6.2f %6d" % (f[:60], s, 100.0 * s / tot, execd[f], 100.0 * execd[f] / te, ninst[f]))
print("total samples", tot, "warp-instructions", te, "cycles/instr (1 warp per scheduler)", round(tot / te * 1.0, 3))
