"""Print an ncu gpu__time_duration launch list in launch order: index, kernel, grid, us (development helper)."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
i = 0
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    i += 1
    if i <= skip:
        continue
    name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void <unnamed>::", "").replace("<unnamed>::", "")
    print(f"{i:5d} {name[:60]:60s} grid={d.get('Grid Size','?'):>14s} {v:9.1f} us")
