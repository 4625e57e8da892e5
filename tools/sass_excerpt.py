"""Static SASS listing of the hot kernels from the built library (cuobjdump -sass; no GPU needed): per kernel the register count,
the opcode histogram and the listing itself (control words dropped) -> profiles/<prefix>_sass_<kernel>.txt.
  python tools/sass_excerpt.py profiles/r02"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kalibr_b200", "libkalibr_b200.so")
prefix = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02")
# (file tag, substring of the mangled name)
KERNELS = [("linearise_assemble_pinhole_radtan", "linearise_assemble_kernelILi0ELb1ELb0E"), ("linearise_materialise_pinhole_radtan", "linearise_materialise_kernelILi0ELb0E"),
           ("set_reduce", "set_reduce_kernel"), ("schur_16x7", "schur_kernelILi16ELi7ELb0E"), ("schur_blocked", "schur_kernelILi8ELi1ELb1E"), ("reduced_solve", "reduced_solve_kernel"), ("sym_eig", "sym_eig_kernel")]
text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", text)
for tag, key in KERNELS:
    body = next((f for f in funcs if f.split("\n", 1)[0].find(key) >= 0), None)
    if body is None:
        print("not found:", key)
        continue
    name = body.split("\n", 1)[0].strip()
    lines = [re.sub(r"\s+", " ", re.sub(r"/\* 0x[0-9a-f]+ \*/", "", ln)).rstrip() for ln in body.split("\n")[1:]]
    ins = [ln for ln in lines if re.search(r"/\*[0-9a-f]{4}\*/", ln)]
    ops = collections.Counter()
    for ln in ins:
        t = re.sub(r"/\*[0-9a-f]{4}\*/", "", ln).split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        ops[op.rstrip(";")] += 1
    usage = ""
    m = re.search(r"Function " + re.escape(name) + r":\s*\n\s*(.*)", res)
    if m:
        usage = m.group(1).strip()
    out = f"{prefix}_sass_{tag}.txt"
    with open(out, "w") as f:
        f.write(f"# cuobjdump -sass kalibr_b200/libkalibr_b200.so (sm_100a), function {name}\n# {usage}\n# {len(ins)} SASS instructions; opcode histogram (static):\n")
        for op, c in ops.most_common(25):
            f.write(f"#   {op:28s} {c}\n")
        f.write("\n".join(ln for ln in ins) + "\n")
    print(tag, len(ins), "instructions;", "DMMA", sum(c for o, c in ops.items() if o.startswith("DMMA")), "|", usage[:100])
