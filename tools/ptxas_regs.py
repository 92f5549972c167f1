"""Register / spill table of every kernel from an `nvcc -Xptxas -v` log (offline check before spending GPU time)."""
import re
import subprocess
import sys

txt = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
blocks = re.split(r"ptxas info\s+: Compiling entry function '", txt)[1:]
rows = []
for b in blocks:
    name = b.split("'", 1)[0]
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
    r = re.search(r"Used (\d+) registers", b)
    rows.append((name, int(r.group(1)), int(m.group(1)), int(m.group(2)), int(m.group(3))))
dem = subprocess.run(["cu++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
for (name, regs, stack, ss, sl), d in zip(rows, dem):
    d = d.replace("vdf::", "")
    m = re.search(r"functor_kernel<(\d+), (\d+), (.*?)>\((?:.*)\)$", d)
    label = f"{m.group(3)} [block {m.group(1)}, minb {m.group(2)}]" if m else d[:120]
    if flt and not re.search(flt, label):
        continue
    print(f"{regs:4d} regs  stack {stack:4d}  spill st/ld {ss}/{sl}   {label}")
