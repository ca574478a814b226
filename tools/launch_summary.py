"""Per-kernel totals of an ncu `--metrics gpu__time_duration.sum --csv` launch list.  python tools/launch_summary.py file.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    name = r[hdr.index("Kernel Name")].split("(")[0][-70:]
    grid = r[hdr.index("Grid Size")]
    val, unit = float(r[hdr.index("Metric Value")]), r[hdr.index("Metric Unit")]
    us = val / 1000 if unit in ("ns", "nsecond") else val
    a = agg.setdefault((name, grid), [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[1]}: one C2 step (virtual splice), ncu gpu__time_duration per launch -- cold-cache, serialised (no PDL overlap):")
print(f"# the SHARE of each kernel is what carries over to the event-timed step, not the absolute sum ({tot:.1f} us)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[0]:4d} x {a[1] / a[0]:8.2f} us = {a[1]:9.1f} us ({100 * a[1] / tot:5.1f}%)  {k[0]} grid {k[1]}")
