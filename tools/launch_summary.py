"""Per-kernel totals of the LAST eager step in an ncu launch list (gpu__time_duration.sum CSV): python tools/launch_summary.py x.csv
A step ends with head_loss_finalize_kernel; times are cold-cache and serialised under ncu, so compare SHARES, not absolutes."""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((r["Kernel Name"], us))
ends = [i for i, (k, _) in enumerate(rows) if "head_loss_finalize" in k]
step = rows[ends[-2] + 1:ends[-1] + 1] if len(ends) >= 2 else rows
tot = sum(us for _, us in step)
agg = collections.OrderedDict()
for k, us in step:
    k = re.sub(r"\(.*", "", k)
    a = agg.setdefault(k, [0.0, 0])
    a[0] += us
    a[1] += 1
print(f"# last eager step of `python bench.py --steps 2 --warmup 3 --eager` under ncu (cold-cache, serialised: compare SHARES)")
print(f"# total {tot:.1f} us over {len(step)} launches")
for k, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{us:9.1f} us {100 * us / tot:5.1f}%  x{n:2d}  {k[:90]}")
