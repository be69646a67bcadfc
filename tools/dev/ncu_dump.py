# dump one device function's instructions (address order) with executed counts and samples from an ncu source csv
#   python tools/dev/ncu_dump.py <source.csv> <kernel-fragment> <function-fragment>
import csv, sys, re, subprocess
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if r[ix["# Samples"]].isdigit()]
base = int(data[0][ix["Address"]], 16)
out = subprocess.run(["cuobjdump", "-elf", "trajectory_planner_b200/csrc/build/tp_vigo.o"], capture_output=True, text=True).stdout
for l in out.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+|0)\s+0x2\s+\S+\s+\S+\s+\$(\S+?)\$(\S+)", l)
    if m and sys.argv[2] in m.group(3) and sys.argv[3] in m.group(4):
        o, s = int(m.group(1), 16), int(m.group(2), 16)
        for r in data:
            a = int(r[ix["Address"]], 16) - base
            if o <= a < o + s:
                print("%5x %9s %6s  %s" % (a - o, r[ix["Instructions Executed"]], r[ix["# Samples"]], r[ix["Source"]].strip()))
