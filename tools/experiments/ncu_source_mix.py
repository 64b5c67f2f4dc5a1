"""Opcode mix and top stall lines from `ncu -i X.ncu-rep --page source --csv ... > file.csv`."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Source" in r)
data = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr)]
iS, iE, iSrc = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"), hdr.index("Source")
num = lambda v: int(v) if v.strip().isdigit() else 0
tot_s, tot_e = sum(num(r[iS]) for r in data), sum(num(r[iE]) for r in data)
print("total samples", tot_s, "warp instructions", tot_e, "static", len(data))
ops, opss = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iSrc])
    op = m.group(2).split(".")[0] if m else "?"
    ops[op] += num(r[iE])
    opss[op] += num(r[iS])
for op, c in ops.most_common(24):
    print(f"{op:12s} inst {100 * c / tot_e:5.1f}%  samples {100 * opss[op] / tot_s:5.1f}%")
print("top stall lines:")
for r in sorted(data, key=lambda r: -num(r[iS]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print(r[iS].rjust(7), r[iE].rjust(9), r[iSrc][:100])
