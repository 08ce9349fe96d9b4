"""Per-launch table from an .ncu-rep: duration, DRAM/L2/SM throughput %, DRAM bytes, issue activity (development helper)."""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
def col(r, name, default=""):
    return r[ix[name]] if name in ix else default
want = [("gpu__time_duration.sum", "us"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB")]
units = rows[1]
print(f"{'#':>3s} {'kernel':44s} " + " ".join(f"{n:>8s}" for _, n in want) + "   GB/s")
for k, r in enumerate(rows[2:]):
    name = re.sub(r"\(.*", "", col(r, "Kernel Name")).replace("void <unnamed>::", "").replace("<unnamed>::", "")[:44]
    vals = []
    for m, n in want:
        v = col(r, m, "nan").replace(",", "")
        try: v = float(v)
        except ValueError: v = float("nan")
        u = units[ix[m]] if m in ix else ""
        if n == "us": v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        if n in ("rdMB", "wrMB"): v = {"byte": v / 1e6, "Kbyte": v / 1e3, "Mbyte": v, "Gbyte": v * 1e3}.get(u, v)
        vals.append(v)
    gbs = (vals[7] + vals[8]) / vals[0] * 1e3 if vals[0] else 0
    print(f"{k:3d} {name:44s} " + " ".join(f"{v:8.1f}" for v in vals) + f" {gbs:7.0f}")
