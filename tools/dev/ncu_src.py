# Summarise an `ncu --page source --csv` export: stall-reason totals and the top instructions by stall samples.
#   python tools/dev/ncu_src.py <source.csv> [topN]
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = 0
items = []
for r in data:
    try:
        s = int(r[ix["# Samples"]])
    except Exception:
        continue
    tot += s
    items.append((s, r))
print("total samples", tot, "instructions", len(items))
agg = collections.Counter()
for s, r in items:
    for c in stall_cols:
        agg[c] += int(r[ix[c]] or 0)
print("stall totals:", [(k, v, round(100.0 * v / max(tot, 1), 1)) for k, v in agg.most_common(10)])
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for rank, (s, r) in enumerate(sorted(items, key=lambda t: -t[0])[:top]):
    st = sorted(((int(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print("%6d %5.2f%%  %-8s exec %-9s %s   [%s]" % (s, 100.0 * s / tot, r[ix["Address"]][-6:], r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:90], ", ".join("%s %d" % (c[6:], v) for v, c in st)))
