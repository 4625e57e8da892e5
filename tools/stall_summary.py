"""Warp-stall samples of one kernel of an ncu report (captured with --set full --import-source on), aggregated by stall reason and by
opcode, plus the executed instruction mix.   python tools/stall_summary.py report.ncu-rep <kernel-regex> [units-divisor]"""
import collections
import csv
import io
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
div = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
seen, uniq = set(), []
for r in data:  # the export lists every instruction twice
    if r[0] in seen:
        continue
    seen.add(r[0])
    uniq.append(r)
data = uniq
smp, src, ex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[smp]) for r in data)
by, cnt, cross, bystall = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
for r in data:
    t = r[src].strip().split()
    o = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    by[o] += int(r[smp])
    cnt[o] += int(r[ex])
    for i in stall_cols:
        v = int(r[i])
        if v:
            bystall[hdr[i][6:]] += v
            cross[o][hdr[i][6:]] += v
print(f"{len(data)} instructions, {tot} samples, {sum(cnt.values()) / div:.1f} executed warp instructions per unit")
print("by stall reason (% of samples):", ", ".join(f"{k} {100 * v / tot:.1f}" for k, v in bystall.most_common(10)))
for o, n in by.most_common(14):
    print(f"  {o:10s} {100 * n / tot:5.1f}%  executed {cnt[o] / div:7.1f}/unit   " + ", ".join(f"{k} {100 * v / tot:.1f}" for k, v in cross[o].most_common(4)))
