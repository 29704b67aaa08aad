"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> markdown: per-kernel launches / time / share per step,
then one step in launch order.  Usage: python tests/tools/launch_list_md.py launches.csv steps > out.md"""
import csv
import re
import sys
from collections import OrderedDict

path, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*$", "", r["Kernel Name"]).replace("void ", "").strip()
    rows.append((name, us, r.get("Grid Size", "")))
# whole steps only: a step starts at vox_insert_kernel
starts = [i for i, r in enumerate(rows) if r[0].startswith("vox_insert_kernel")]
if len(starts) >= 2:
    rows_use = rows[starts[0]:starts[-1]]
    nsteps = len(starts) - 1
else:
    rows_use, nsteps = rows, steps
agg = OrderedDict()
for n, us, g in rows_use:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(v[1] for v in agg.values())
print("| kernel | launches/step | us/step | share | us each |")
print("|---|---:|---:|---:|---:|")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %.1f | %.1f | %.1f %% | %.1f |" % (n, c / nsteps, t / nsteps, 100 * t / tot, t / c))
print("\nTotal %.1f us of kernel time per step (%d launches, %d complete steps averaged)." % (tot / nsteps, len(rows_use) // nsteps, nsteps))
print("\nOne step in launch order (us):\n\n```")
for n, us, g in rows_use[: len(rows_use) // nsteps]:
    print("%-55s %6.1f  grid %s" % (n[:55], us, g))
print("```")
