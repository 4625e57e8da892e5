"""Summarise `ncu -i X.ncu-rep --page source --csv` (SASS view): share of stall samples and executed instructions per opcode,
stall reasons, and the hottest instructions."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0  # n-th kernel of the report
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
rows = rows[starts[which]:starts[which + 1]]
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
isrc, ismp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ismp]) for r in data)
totex = sum(int(r[iex]) for r in data)
print("kernel:", rows[0][1][:100] if rows[0][0] == "Kernel Name" else "?")
print("samples", tot, "warp instructions", totex, "SASS lines", len(data))
cat, catex = collections.Counter(), collections.Counter()
for r in data:
    t = r[isrc].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    cat[op] += int(r[ismp]); catex[op] += int(r[iex])
print("| opcode | samples % | executed % |\n|---|---:|---:|")
for op, c in cat.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 18):
    print(f"| {op} | {100 * c / tot:.1f} | {100 * catex[op] / totex:.1f} |")
st = collections.Counter()
for r in data:
    for i, h in stalls:
        st[h] += int(r[i] or 0)
print("stall reasons (% of samples):", {k: round(100 * v / tot, 1) for k, v in st.most_common(8)})
print("hottest instructions:")
for r in sorted(data, key=lambda r: -int(r[ismp]))[:12]:
    print(f"  {100 * int(r[ismp]) / tot:5.2f}%  {r[isrc][:110]}")
