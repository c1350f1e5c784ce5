"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i
        break
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    agg.setdefault(r[ki][:72], []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print("%-72s n=%3d  %10.3f ms  %5.1f%%" % (k, len(v), sum(v) / 1e6, 100 * sum(v) / tot))
print("total %.3f ms" % (tot / 1e6))
