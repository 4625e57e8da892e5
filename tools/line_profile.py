"""Map the stall samples of an `ncu --page source --csv` SASS export to CUDA source lines using `nvdisasm -g -c` output
of the same cubin (instruction i of the export = instruction i of the disassembly).
  python tools/line_profile.py sass.csv dis.txt <mangled-name-substring> [kernel-index-in-export] [top-n]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
rows = rows[starts[which]:starts[which + 1]]
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
ismp, iex, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
lines = open(sys.argv[2]).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and sys.argv[3] in l][0]
cur, seq = None, []
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]+\*/", l):
        seq.append((cur, l.split("*/", 1)[1].strip()))
assert abs(len(seq) - len(data)) <= 2, (len(seq), len(data))
agg, aggex = collections.Counter(), collections.Counter()
for (ln, _), r in zip(seq, data):
    agg[ln] += int(r[ismp]); aggex[ln] += int(r[iex])
tot, totex = sum(agg.values()), sum(aggex.values())
import os
srcs = {}
def text(key):
    if not key: return ""
    f, ln = key
    if f not in srcs:
        pth = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "kalibr_b200", "csrc", f)
        srcs[f] = open(pth).read().split("\n") if os.path.exists(pth) else []
    return srcs[f][ln - 1].strip()[:100] if ln - 1 < len(srcs[f]) else ""
print(f"samples {tot}, instructions {totex}")
for ln, c in agg.most_common(int(sys.argv[5]) if len(sys.argv) > 5 else 40):
    print(f"{100 * c / tot:5.1f}% smp {100 * aggex[ln] / totex:5.1f}% ins  {ln[0] if ln else '?'}:{ln[1] if ln else 0}: {text(ln)}")
