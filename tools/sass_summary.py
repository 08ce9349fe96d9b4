"""SASS evidence for profiles/: per kernel of csrc/libffb200.so the counts of the Blackwell-specific mnemonics
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops) next to HMMA (mma.sync) and MUFU.  Run in the build container (no GPU needed):
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "image-super-resolution-2_b200", "csrc", "libffb200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
want = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "MUFU.EX2", "MUFU.TANH", "LDGSTS", "FFMA2", "FADD2"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("void ", "")
        name = re.sub(r"\((?!int\)).*", "", name).replace("(int)", "")
        cur = name
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_all"] += 1
        for w in want:
            if op.startswith(w):
                counts[cur][w] += 1
                total[w] += 1
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)} (sm_100a): instruction counts per kernel; columns: " + " ".join(want))
print(f"{'kernel':86s} {'instr':>7s} " + " ".join(f"{w[:8]:>8s}" for w in want))
for k, c in counts.items():
    if not any(c[w] for w in want[:8]):
        continue
    print(f"{k[:86]:86s} {c['_all']:7d} " + " ".join(f"{c[w]:8d}" for w in want))
print(f"{'TOTAL':86s} {sum(c['_all'] for c in counts.values()):7d} " + " ".join(f"{total[w]:8d}" for w in want))
hm = [k for k, c in counts.items() if c["HMMA"]]
print("# kernels that contain HMMA (mma.sync):", ", ".join(sorted(set(re.sub(r"<.*", "", k) for k in hm))) or "none")
