"""Attribute the stall samples of an .ncu-rep to source lines (read on the CPU box).

ncu's CLI source page only prints SASS; this joins it, instruction by instruction, with `nvdisasm -g` line info of
the same kernel in the in-tree library (so the library must be the build that was profiled).
usage: python tools/ncu_lines.py gpurun_out/clip_full.ncu-rep clip_kernelILb1ELb0E [n_top] [exec]
columns: share of stall samples, share of warp instructions executed, warp instructions, source line"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, fun = sys.argv[1], sys.argv[2]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 40
by_exec = len(sys.argv) > 4 and sys.argv[4] == "exec"      # rank by warp instructions executed instead of stall samples

sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
h = rows[1]
ia, isrc, isamp, iexe = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
prof, seen = [], set()
for r in rows[2:]:
    if len(r) <= isamp or not r[isamp].isdigit() or r[ia] in seen:
        continue
    seen.add(r[ia])
    prof.append((r[isrc].strip(), int(r[isamp]), int(r[iexe])))

with tempfile.TemporaryDirectory(dir=os.path.join(ROOT, "gpurun_out")) as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "diffusion_conductor_b200", "libdc_b200.so")], cwd=td, capture_output=True)
    dis = subprocess.run(["nvdisasm", "-g", "-c", glob.glob(os.path.join(td, "*.cubin"))[0]], capture_output=True, text=True).stdout
lines, on, cur = [], False, ("?", 0)
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        on = fun in ln
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if m:
        lines.append((cur, m.group(2).strip()))
if len(lines) != len(prof):
    print(f"warning: {len(lines)} disassembled vs {len(prof)} profiled instructions; joining the common prefix")
by_line, exe_line = collections.Counter(), collections.Counter()
for (loc, _), (_, s, e) in zip(lines, prof):
    by_line[loc] += s
    exe_line[loc] += e
tot = sum(by_line.values()) or 1
src = {}
print(f"{tot} samples over {len(prof)} instructions; top source lines:")
tot_e = sum(exe_line.values()) or 1
order = exe_line.most_common(ntop) if by_exec else by_line.most_common(ntop)
for (f, n), _ in order:
    s = by_line[(f, n)]
    if f not in src:
        cand = glob.glob(os.path.join(ROOT, "diffusion_conductor_b200", "csrc", f))
        src[f] = open(cand[0]).read().splitlines() if cand else []
    text = src[f][n - 1].strip()[:110] if 0 < n <= len(src[f]) else ""
    print(f"{s / tot:6.1%} {exe_line[(f, n)] / tot_e:6.1%} {exe_line[(f, n)]:>9} {f}:{n}  {text}")
