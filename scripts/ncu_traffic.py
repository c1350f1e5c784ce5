"""Extract dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full reports and
write profiles/ncu_traffic.json (read by bench.py for roofline.traffic).
usage: python scripts/ncu_traffic.py name=report.ncu-rep:units_in_that_launch ..."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
SC = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
for arg in sys.argv[1:]:
    name, rest = arg.split("=")
    rep, units = rest.rsplit(":", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, un = rows[0], rows[1]
    for r in rows[2:]:
        if name.split("@")[0] not in r[hdr.index("Kernel Name")]:
            continue
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            tot += float(r[i].replace(",", "")) * SC[un[i]]
        out[name] = {"report": os.path.basename(rep), "units_in_launch": float(units), "dram_bytes": tot,
                     "dram_bytes_per_unit": tot / float(units),
                     "gpu_time_us": float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")) *
                     {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}[un[hdr.index("gpu__time_duration.sum")]]}
        break
json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
