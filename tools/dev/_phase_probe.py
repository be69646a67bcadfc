# aggregate the per-trajectory phase cycle counts printed by the TP_LBFGS_TIMING build (tools: build_timing.sh)
import os, re, subprocess, sys
env = dict(os.environ, PROBE_B=os.environ.get("PROBE_B", "4096"))
out = subprocess.run([sys.executable, "/root/repo/tools/_solve_probe.py"], env=env, capture_output=True, text=True).stdout
tot = [0, 0, 0, 0, 0]; n = 0
for m in re.finditer(r"total (\d+) init (\d+) lbfgs (\d+) collision (\d+) step (\d+)", out):
    v = list(map(int, m.groups())); n += 1
    for i in range(5): tot[i] += v[i]
print("trajectories", n, "kcycles: total %d init %d (%.1f%%) lbfgs %d (%.1f%%) collision %d (%.1f%%) step %d (%.1f%%)" % (
    tot[0], tot[1], 100 * tot[1] / tot[0], tot[2], 100 * tot[2] / tot[0], tot[3], 100 * tot[3] / tot[0], tot[4], 100 * tot[4] / tot[0]))
