"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export by CUDA source line.
usage: python scripts/ncu_lines.py file.csv [ntop]"""
import csv, sys, collections
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(open(sys.argv[1])))
agg = collections.Counter(); inst = collections.Counter(); text = {}; stall = collections.defaultdict(collections.Counter)
fname = ""; hdr = None; cur = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ci = {}; [ci.setdefault(h, i) for i, h in enumerate(hdr)]; continue
    if r[0] == "Function Name" or hdr is None: continue
    if r[0].strip():
        cur = (fname, int(r[0])); text[cur] = r[1].strip()
    try:
        s = int(r[ci["# Samples"]] or 0)
    except (ValueError, IndexError):
        continue
    agg[cur] += s; inst[cur] += int(r[ci["Instructions Executed"]] or 0)
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stall[cur][h[6:]] += int(r[ci[h]] or 0)
tot = sum(agg.values())
print("total samples", tot)
for k, v in agg.most_common(ntop):
    top = " ".join("%s=%d" % kv for kv in stall[k].most_common(2))
    print("%6d %5.1f%% %10d  %s:%d  %-70s %s" % (v, 100.0 * v / tot, inst[k], k[0], k[1], text[k][:70], top))
