"""Mean DRAM bytes per launch of the conv_gemm launches in an `ncu --set full` report -> profiles/*_conv_gemm_traffic.json."""
import csv, io, json, subprocess, sys
rep, out, source = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot, n, per = 0.0, 0, []
for r in rows[2:]:
    if "conv_gemm_tc_kernel" not in r[ix["Kernel Name"]]:
        continue
    b = sum(float(r[ix[m]].replace(",", "")) * scale.get(units[ix[m]], 1.0) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    us = float(r[ix["gpu__time_duration.sum"]].replace(",", "")) / (1e3 if units[ix["gpu__time_duration.sum"]] in ("ns", "nsecond") else 1.0)
    per.append({"kernel": r[ix["Kernel Name"]][:60], "us": us, "dram_bytes": b})
    tot += b; n += 1
json.dump({"source": source, "mean_dram_bytes_per_launch": tot / max(n, 1), "launches": n, "per_launch": per}, open(out, "w"), indent=1)
print(out, n, tot / max(n, 1))
