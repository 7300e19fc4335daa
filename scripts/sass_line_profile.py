"""
Source-line profile of one kernel from an `ncu --page source --csv` export (SASS rows with stall
samples) joined with `nvdisasm -g` line information of the same build:

  cuobjdump -xelf all fitburst_b200/libfitburst_b200.so            (in a scratch directory)
  nvdisasm -g -c fb_api.sm_100a.cubin > all.sass
  python scripts/sass_line_profile.py all.sass _ZN2fb7k_vals3ILi8EEEvNS_8EvalArgsEPd source.csv [kernel instance]

Prints the source lines with the most warp-stall samples and their share of executed instructions.
"""
import collections
import csv
import os
import re
import sys

sass_path, symbol, csv_path = sys.argv[1:4]
instance = int(sys.argv[4]) if len(sys.argv) > 4 else 1
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

inside, cur, off2line = False, None, {}
for line in open(sass_path):
    if line.startswith("\t.section\t.text."):
        inside = (".text." + symbol) in line
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", line)
    if m:
        off2line[int(m.group(1), 16)] = cur

head, seen, base = None, 0, None
samples, insts, by_op = collections.Counter(), collections.Counter(), collections.Counter()
for row in csv.reader(open(csv_path)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        seen += 1
        continue
    if row[0] == "Address":
        head = row
        si, ie, so = head.index("# Samples"), head.index("Instructions Executed"), head.index("Source")
        continue
    if head is None or seen != instance:
        continue
    try:
        addr, s, n = int(row[0], 16), float(row[si]), float(row[ie])
    except ValueError:
        continue
    base = addr if base is None else base
    key = off2line.get(addr - base) or ("?", 0)
    samples[key] += s
    insts[key] += n
    op = row[so].strip().split()
    by_op[(op[1] if op and op[0].startswith("@") else (op[0] if op else "?")).split(".")[0]] += n
total_s, total_n = sum(samples.values()), sum(insts.values())
print(f"{symbol}: {total_n:.3g} warp instructions, {total_s:.0f} stall samples")
print("opcode mix (share of executed warp instructions):",
      ", ".join(f"{op} {100 * n / total_n:.1f}%" for op, n in by_op.most_common(14)))
sources = {}
for (name, line_no), s in samples.most_common(30):
    if name not in sources:
        path = os.path.join(root, "fitburst_b200", "csrc", name)
        sources[name] = open(path).read().split("\n") if os.path.exists(path) else []
    text = sources[name][line_no - 1].strip()[:100] if 0 < line_no <= len(sources[name]) else ""
    print(f"{name}:{line_no:5d}  samples {100 * s / total_s:5.1f}%  instr {100 * insts[(name, line_no)] / total_n:5.1f}%  {text}")
