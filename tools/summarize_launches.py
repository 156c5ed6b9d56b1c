"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/summarize_launches.py
launches.csv [first_launch_id_of_the_step]  (default: the last training step of the list)."""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = val / 1000.0 if unit.startswith("n") else (val if unit.startswith("u") else val * 1000.0)
    rows.append((int(r["ID"]), re.sub(r"\(.*", "", r["Kernel Name"]).replace("hrnb::", ""), us))
rows.sort()
if len(sys.argv) > 2:
    start = int(sys.argv[2])
else:       # the last training step of the list starts with three zero-fills followed by the stem im2col kernel
    stems = [i for i, name, _ in rows if "stem_im2col" in name]
    start = stems[-1] - 3 if stems else rows[len(rows) // 2][0]
sel = [r for r in rows if r[0] >= start]
agg = collections.defaultdict(lambda: [0, 0.0])
for _, name, us in sel:
    name = re.sub(r"<.*", "", name)
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-42s n=%4d %9.1f us %5.1f %%" % (name, n, us, 100.0 * us / tot))
print("total %d launches %.1f us" % (len(sel), tot))
