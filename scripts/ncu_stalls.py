"""Summarise an `ncu --page source --csv` export: stall-reason totals and the hottest SASS lines.
usage: python scripts/ncu_stalls.py file.csv [ntop]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); total_samples = 0
lines = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    s = int(r[ci["# Samples"]] or 0); total_samples += s
    for h in stall_cols: tot[h] += int(r[ci[h]] or 0)
    lines.append((s, r[ci["Source"]].strip(), int(r[ci["Instructions Executed"]] or 0), r))
print("total samples", total_samples)
for h, v in tot.most_common(12): print("  %-28s %8d %5.1f%%" % (h, v, 100.0 * v / max(total_samples, 1)))
ops = collections.Counter(); opsamp = collections.Counter()
for s, src, ie, r in lines:
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    ops[op] += ie; opsamp[op] += s
print("instruction mix (warp-level executed, samples):")
for op, v in ops.most_common(18): print("  %-14s %12d  samples %7d" % (op, v, opsamp[op]))
print("hottest lines:")
for i, (s, src, ie, r) in enumerate(sorted(enumerate(lines), key=lambda x: -x[1][0])[:ntop] and sorted(lines, key=lambda x: -x[0])[:ntop]):
    top = sorted(((int(r[ci[h]] or 0), h) for h in stall_cols), reverse=True)[:2]
    print("  %6d  %-70s %s" % (s, src[:70], " ".join("%s=%d" % (h[6:], v) for v, h in top)))
