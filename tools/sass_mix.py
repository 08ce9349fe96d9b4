"""Instruction mix / shared-memory wavefronts / stall reasons per SASS opcode from `ncu --page source --csv` (development helper)."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = None
blocks = []
for r in rows:
    if "Source" in r and "Address" in r:
        hdr = r
        blocks.append([])
        continue
    if hdr and len(r) == len(hdr):
        blocks[-1].append(r)
i = {h: k for k, h in enumerate(hdr)}
body = blocks[want]
def num(s):
    try: return int(float(s))
    except Exception: return 0
mix, samples, wf, wfi, stall = (collections.Counter() for _ in range(5))
tot = 0
for r in body:
    src = r[i["Source"]]
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = ".".join((m.group(2) if m else src[:10]).split(".")[:2])
    n = num(r[i["Instructions Executed"]])
    mix[op] += n; tot += n
    samples[op] += num(r[i["# Samples"]])
    wf[op] += num(r[i["L1 Wavefronts Shared"]]); wfi[op] += num(r[i["L1 Wavefronts Shared Ideal"]])
    for s in hdr:
        if s.startswith("stall_") and "Not Issued" not in s:
            stall[s] += num(r[i[s]])
print("blocks", len(blocks), "total warp-instructions", tot)
for op, n in mix.most_common(28):
    print(f"{op:20s} {n:11d} {100*n/max(tot,1):5.1f}%  samples {samples[op]:7d}  smem wavefronts {wf[op]} (ideal {wfi[op]})")
print({k: v for k, v in stall.most_common()})
