"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", "")) * {"us": 1e-3, "ns": 1e-6, "ms": 1.0}.get(r[ui], 1.0)
    agg.setdefault(r[ki].split("(")[0], []).append(v)
skip = sys.argv[2:] 
tot = sum(sum(v) for k, v in agg.items() if not any(s in k for s in skip))
print("| kernel | launches | avg ms | share |\n|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    sh = "-" if any(s in k for s in skip) else f"{100 * sum(v) / tot:.1f}%"
    print(f"| `{k}` | {len(v)} | {sum(v) / len(v):.4f} | {sh} |")
