# sizes (bytes) of the device functions inside one kernel of tp_vigo.o:  python tools/dev/fsizes.py k_solve_w
import subprocess, sys, re
pat = sys.argv[1]
obj = sys.argv[2] if len(sys.argv) > 2 else "trajectory_planner_b200/csrc/build/tp_vigo.o"
out = subprocess.run(["cuobjdump", "-elf", obj], capture_output=True, text=True).stdout
rows = []
for l in out.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+|0)\s+0x2\s+\S+\s+\S+\s+\$(\S+?)\$(\S+)", l)
    if m and pat in m.group(3):
        rows.append((int(m.group(2), 16), m.group(4)))
for s, n in sorted(rows):
    print("%7d  %s" % (s, n))
print("total", sum(s for s, _ in rows))
