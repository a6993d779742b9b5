"""ncu launch list (`--metrics gpu__time_duration.sum --csv`) -> profiles/<name>: launch_index,kernel,duration_us and
the share of each of the library's kernels in the listed time.  usage: launch_list.py raw.csv out.csv"""
import csv
import sys
from collections import defaultdict

raw, out = sys.argv[1], sys.argv[2]
rows = []
with open(raw) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    rows.append((int(r["ID"]), r["Kernel Name"][:90], us))
with open(out, "w") as f:
    f.write("launch_index,kernel,duration_us\n")
    for i, k, us in rows:
        f.write(f'{i},"{k}",{us:.2f}\n')
tot = defaultdict(lambda: [0, 0.0])
for _, k, us in rows:
    if "plk::" in k or "infonce" in k or "l2norm" in k or "grad_finish" in k or "loss_kernel" in k:
        name = k.split("(")[0].replace("void ", "").replace("plk::", "")
        tot[name][0] += 1
        tot[name][1] += us
s = sum(v[1] for v in tot.values())
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:50s} {n:4d} launches  avg {us / n:8.2f} us  {100 * us / s:5.1f} % of the library's listed time")
